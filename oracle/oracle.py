"""ctypes binding of the CPU oracle (oracle/libmiqp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from .dat_io import FlatProblem

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmiqp_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class OrcProblem(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("N", "R", "C", "O", "L", "E")]
        + [(n, C.c_double) for n in (
            "ts", "min_vel", "max_vel", "total_min_acc", "total_max_acc", "total_min_jerk",
            "total_max_jerk", "maximum_slack", "w_slack", "w_slack_obs",
            "min_region_change_speed", "gap_tol", "time_limit")]
        + [("safety", _dp), ("safety_slack", _dp)]
        + [(n, _dp) for n in ("w_pos_x", "w_vel_x", "w_acc_x", "w_pos_y", "w_vel_y", "w_acc_y",
                              "w_jerk_x", "w_jerk_y", "wheelbase", "radius", "x0",
                              "x_ref", "vx_ref", "y_ref", "vy_ref",
                              "min_acc_x", "max_acc_x", "min_acc_y", "max_acc_y",
                              "min_jerk_x", "max_jerk_x", "min_jerk_y", "max_jerk_y")]
        + [("initial_region", _ip), ("possible_region", _ip), ("obs_edges", _dp),
           ("obs_nedges", _ip), ("obs_soft", _ip), ("env_edges", _dp), ("env_off", _ip),
           ("frac", _dp)]
        + [(n, _dp) for n in ("poly_sint_ub", "poly_sint_lb", "poly_coss_ub", "poly_coss_lb",
                              "poly_kappa_max", "poly_kappa_min")]
    )


class OrcSizes(C.Structure):
    _fields_ = [("ncols", C.c_int), ("ncont", C.c_int), ("nbin", C.c_int),
                ("nrows", C.c_long), ("nnz_struct", C.c_long), ("nnz", C.c_long)]


class OrcLayout(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "C", "N", "R", "O", "L", "E", "K", "base_nwe", "base_ar", "base_rcna", "base_dcc",
        "base_dcf", "base_so", "base_sof", "base_c2c", "base_sv", "ncols")]


class OrcSolveInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("objective", C.c_double), ("best_bound", C.c_double),
                ("gap", C.c_double), ("seconds", C.c_double), ("max_violation", C.c_double),
                ("nodes", C.c_long), ("qp_solves", C.c_long), ("qp_iters", C.c_long),
                ("proven", C.c_int), ("uncertified", C.c_long)]


def build(force: bool = False) -> str:
    """Compile the oracle (plain C, gcc).  Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in ("miqp_oracle.c", "miqp_oracle_bnb.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if not force and os.path.exists(_LIB_PATH):
        newest = max(os.path.getmtime(s) for s in srcs + [os.path.join(_HERE, "miqp_oracle.h")])
        if os.path.getmtime(_LIB_PATH) >= newest:
            return _LIB_PATH
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
           "-o", _LIB_PATH] + srcs + ["-lm", "-lpthread"]
    subprocess.check_call(cmd)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_objective.restype = C.c_double
        _lib.orc_max_violation.restype = C.c_double
        _lib.orc_build_rows.restype = C.c_long
        if hasattr(_lib, "orc_complete_assignment"):
            _lib.orc_complete_assignment.restype = C.c_double
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


@dataclass
class COracleProblem:
    struct: OrcProblem
    keep: list


def to_c(p: FlatProblem, gap_tol: float | None = None, time_limit: float | None = None) -> COracleProblem:
    s = OrcProblem()
    keep = []
    s.N, s.R, s.C, s.O, s.L, s.E = p.N, p.R, p.C, p.O, p.L, p.E
    sc = p.scal
    s.ts = sc["ts"]
    s.min_vel, s.max_vel = sc["min_vel_x_y"], sc["max_vel_x_y"]
    s.total_min_acc, s.total_max_acc = sc["total_min_acc"], sc["total_max_acc"]
    s.total_min_jerk, s.total_max_jerk = sc["total_min_jerk"], sc["total_max_jerk"]
    s.maximum_slack = sc["maximum_slack"]
    s.w_slack, s.w_slack_obs = sc["WEIGHTS_SLACK"], sc["WEIGHTS_SLACK_OBSTACLE"]
    s.min_region_change_speed = sc["minimum_region_change_speed"]
    s.gap_tol = sc["relative_mip_gap_tolerance"] if gap_tol is None else gap_tol
    s.time_limit = sc["max_solution_time"] if time_limit is None else time_limit

    def setd(name, arr):
        a, ptr = _d(arr)
        keep.append(a)
        setattr(s, name, ptr)

    def seti(name, arr):
        a, ptr = _i(arr)
        keep.append(a)
        setattr(s, name, ptr)

    setd("safety", p.safety)
    setd("safety_slack", p.safety_slack)
    for cn, key in (("w_pos_x", "WEIGHTS_POS_X"), ("w_vel_x", "WEIGHTS_VEL_X"), ("w_acc_x", "WEIGHTS_ACC_X"),
                    ("w_pos_y", "WEIGHTS_POS_Y"), ("w_vel_y", "WEIGHTS_VEL_Y"), ("w_acc_y", "WEIGHTS_ACC_Y"),
                    ("w_jerk_x", "WEIGHTS_JERK_X"), ("w_jerk_y", "WEIGHTS_JERK_Y"),
                    ("wheelbase", "WheelBase"), ("radius", "CollisionRadius")):
        setd(cn, p.car[key])
    setd("x0", p.x0)
    for k in ("x_ref", "vx_ref", "y_ref", "vy_ref"):
        setd(k, p.ref[k])
    for k in ("min_acc_x", "max_acc_x", "min_acc_y", "max_acc_y",
              "min_jerk_x", "max_jerk_x", "min_jerk_y", "max_jerk_y"):
        setd(k, p.lim[k])
    seti("initial_region", p.initial_region)
    seti("possible_region", p.possible_region)
    setd("obs_edges", p.obs_edges if p.obs_edges.size else np.zeros(4))
    seti("obs_nedges", p.obs_nedges if p.obs_nedges.size else np.zeros(1, dtype=np.int32))
    seti("obs_soft", p.obs_soft if p.obs_soft.size else np.zeros(1, dtype=np.int32))
    setd("env_edges", p.env_edges if p.env_edges.size else np.zeros(4))
    seti("env_off", p.env_off)
    setd("frac", p.frac)
    for cn, key in (("poly_sint_ub", "POLY_SINT_UB"), ("poly_sint_lb", "POLY_SINT_LB"),
                    ("poly_coss_ub", "POLY_COSS_UB"), ("poly_coss_lb", "POLY_COSS_LB"),
                    ("poly_kappa_max", "POLY_KAPPA_AX_MAX"), ("poly_kappa_min", "POLY_KAPPA_AX_MIN")):
        setd(cn, p.poly[key])
    return COracleProblem(s, keep)


def sizes(p: FlatProblem) -> OrcSizes:
    cp = to_c(p)
    out = OrcSizes()
    lib().orc_sizes(C.byref(cp.struct), C.byref(out))
    return out


def layout(p: FlatProblem) -> OrcLayout:
    cp = to_c(p)
    out = OrcLayout()
    lib().orc_layout(C.byref(cp.struct), C.byref(out))
    return out


def build_rows(p: FlatProblem):
    """Structural CSR (explicit zeros kept) in OPL row order."""
    cp = to_c(p)
    sz = sizes(p)
    rowptr = np.zeros(sz.nrows + 1, dtype=np.int64)
    cols = np.zeros(sz.nnz_struct, dtype=np.int32)
    vals = np.zeros(sz.nnz_struct, dtype=np.float64)
    lo = np.zeros(sz.nrows)
    hi = np.zeros(sz.nrows)
    lib().orc_build_rows(C.byref(cp.struct), rowptr.ctypes.data_as(C.POINTER(C.c_long)),
                         cols.ctypes.data_as(_ip), vals.ctypes.data_as(_dp),
                         lo.ctypes.data_as(_dp), hi.ctypes.data_as(_dp))
    return rowptr, cols, vals, lo, hi


def col_info(p: FlatProblem):
    cp = to_c(p)
    n = layout(p).ncols
    isb = np.zeros(n, dtype=np.uint8)
    lb = np.zeros(n)
    ub = np.zeros(n)
    lib().orc_col_info(C.byref(cp.struct), isb.ctypes.data_as(C.POINTER(C.c_ubyte)),
                       lb.ctypes.data_as(_dp), ub.ctypes.data_as(_dp))
    return isb, lb, ub


def objective(p: FlatProblem, x: np.ndarray) -> float:
    cp = to_c(p)
    x = np.ascontiguousarray(x, dtype=np.float64)
    return float(lib().orc_objective(C.byref(cp.struct), x.ctypes.data_as(_dp)))


def max_violation(p: FlatProblem, x: np.ndarray):
    cp = to_c(p)
    x = np.ascontiguousarray(x, dtype=np.float64)
    worst = C.c_long(-1)
    v = float(lib().orc_max_violation(C.byref(cp.struct), x.ctypes.data_as(_dp), C.byref(worst)))
    return v, worst.value


def complete_assignment(p: FlatProblem, x: np.ndarray):
    cp = to_c(p)
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    v = float(lib().orc_complete_assignment(C.byref(cp.struct), x.ctypes.data_as(_dp)))
    return x, v


def solve(p: FlatProblem, gap_tol: float | None = None, time_limit: float | None = None,
          warm: np.ndarray | None = None, verbose: int = 0):
    cp = to_c(p, gap_tol, time_limit)
    n = layout(p).ncols
    x = np.zeros(n)
    info = OrcSolveInfo()
    wp = None
    if warm is not None:
        warm = np.ascontiguousarray(warm, dtype=np.float64)
        wp = warm.ctypes.data_as(_dp)
    lib().orc_solve(C.byref(cp.struct), wp, x.ctypes.data_as(_dp), C.byref(info), int(verbose))
    return x, info


def solve_fixed(p: FlatProblem, x_bin: np.ndarray):
    cp = to_c(p)
    n = layout(p).ncols
    x = np.zeros(n)
    obj = C.c_double(0.0)
    xb = np.ascontiguousarray(x_bin, dtype=np.float64)
    rc = lib().orc_solve_fixed(C.byref(cp.struct), xb.ctypes.data_as(_dp), x.ctypes.data_as(_dp), C.byref(obj))
    return rc, x, obj.value


# ---- helpers to move between the full column vector and named blocks -------------------
CORE_BLOCKS = ["u_x", "u_y", "pos_x", "vel_x", "acc_x", "pos_y", "vel_y", "acc_y",
               "pos_x_front_UB", "pos_x_front_LB", "pos_y_front_UB", "pos_y_front_LB"]
NWE_NAMES = ["notWithinEnvironmentRear", "notWithinEnvironmentFrontUbUb", "notWithinEnvironmentFrontLbUb",
             "notWithinEnvironmentFrontUbLb", "notWithinEnvironmentFrontLbLb"]
RCNA_NAMES = ["region_change_not_allowed_x_positive", "region_change_not_allowed_y_positive",
              "region_change_not_allowed_x_negative", "region_change_not_allowed_y_negative",
              "region_change_not_allowed_combined"]


def block_views(p: FlatProblem, x: np.ndarray) -> dict:
    """Named views (RawResults families, src/miqp_planner_data.hpp:46-97) into x."""
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
