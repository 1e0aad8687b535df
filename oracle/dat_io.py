"""OPL ``.dat`` reader / writer and the flat problem container used by the tests.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under ``oracle/`` is imported
by the product path.  The product has its own C++ reader
(``planner-miqp_b200/csrc/dat_reader.cpp``); the two are checked against each other.

Grammar follows the two dialects found in the reference fixtures
(``cplexmodel/cplexmodel_testcase.dat`` hand written, ``cplexmodel/test_sos.dat`` OPL
printed): ``name = value;`` where value is a number, a nested
``[...]`` array, a ``{...}`` set of ``<...>`` tuples; separators are whitespace and/or
commas; ``/* */`` and ``//`` comments.

Element names and order follow ``src/model_input_data_source.cpp:180-275`` of the
reference.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any

import numpy as np

_TOKEN = re.compile(
    r"\s*(?:(/\*.*?\*/)|(//[^\n]*)|([\[\]{}<>,;=])|"
    r"([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|[iI]nfinity))|([A-Za-z_][A-Za-z_0-9]*))",
    re.S,
)


def _tokenize(text: str):
    pos = 0
    n = len(text)
    out = []
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise ValueError(f"dat: cannot tokenize at {pos}: {text[pos:pos+40]!r}")
        pos = m.end()
        if m.group(1) or m.group(2):
            continue
        if m.group(3):
            out.append(m.group(3))
        elif m.group(4):
            out.append(("num", m.group(4)))
        else:
            out.append(("id", m.group(5)))
    return out


def _parse_value(tok, i):
    t = tok[i]
    if t == "[":
        i += 1
        items = []
        while tok[i] != "]":
            if tok[i] == ",":
                i += 1
                continue
            v, i = _parse_value(tok, i)
            items.append(v)
        return items, i + 1
    if t == "{":
        i += 1
        items = []
        while tok[i] != "}":
            if tok[i] == ",":
                i += 1
                continue
            v, i = _parse_value(tok, i)
            items.append(v)
        return {"set": items}, i + 1
    if t == "<":
        i += 1
        items = []
        while tok[i] != ">":
            if tok[i] == ",":
                i += 1
                continue
            v, i = _parse_value(tok, i)
            items.append(v)
        return tuple(items), i + 1
    if isinstance(t, tuple) and t[0] == "num":
        s = t[1]
        if re.fullmatch(r"[-+]?\d+", s):
            return int(s), i + 1
        return float(s), i + 1
    raise ValueError(f"dat: unexpected token {t!r}")


def parse_dat_text(text: str) -> dict[str, Any]:
    tok = _tokenize(text)
    i = 0
    out: dict[str, Any] = {}
    while i < len(tok):
        t = tok[i]
        if t == ";":
            i += 1
            continue
        if not (isinstance(t, tuple) and t[0] == "id"):
            raise ValueError(f"dat: expected identifier, got {t!r}")
        name = t[1]
        if tok[i + 1] != "=":
            raise ValueError(f"dat: expected '=' after {name}")
        v, i = _parse_value(tok, i + 2)
        out[name] = v
        if i < len(tok) and tok[i] == ";":
            i += 1
    return out


SOLVER_INT_KEYS = ["mipdisplay", "mipemphasis", "cutpass", "probe", "repairtries",
                   "rinsheur", "varsel", "mircuts", "parallelmode"]
CAR_VEC_KEYS = ["WEIGHTS_POS_X", "WEIGHTS_VEL_X", "WEIGHTS_ACC_X", "WEIGHTS_POS_Y",
                "WEIGHTS_VEL_Y", "WEIGHTS_ACC_Y", "WEIGHTS_JERK_X", "WEIGHTS_JERK_Y",
                "WheelBase", "CollisionRadius"]
CAR_STEP_KEYS = ["x_ref", "vx_ref", "y_ref", "vy_ref"]
CAR_REGION_KEYS = ["min_acc_x", "max_acc_x", "min_acc_y", "max_acc_y",
                   "min_jerk_x", "max_jerk_x", "min_jerk_y", "max_jerk_y"]
POLY_KEYS = ["POLY_SINT_UB", "POLY_SINT_LB", "POLY_COSS_UB", "POLY_COSS_LB",
             "POLY_KAPPA_AX_MAX", "POLY_KAPPA_AX_MIN"]
SCALAR_REAL_KEYS = ["max_solution_time", "relative_mip_gap_tolerance", "relobjdif", "ts",
                    "min_vel_x_y", "max_vel_x_y", "total_min_acc", "total_max_acc",
                    "total_min_jerk", "total_max_jerk", "maximum_slack", "WEIGHTS_SLACK",
                    "WEIGHTS_SLACK_OBSTACLE", "minimum_region_change_speed"]


@dataclass
class FlatProblem:
    """One MIQP instance = the content of the reference's ``ModelParameters``
    (``src/miqp_planner_data.hpp:99-185``) in plain row-major numpy arrays.

    Polygons are stored as closed edge lists ``(x1, y1, x2, y2)`` exactly as
    ``ModelInputDataSource::addLineSet`` (``src/model_input_data_source.cpp:167-178``)
    hands them to OPL: edge k runs from vertex k to vertex k+1, the last edge closes.
    """
    N: int = 0
    R: int = 0
    C: int = 0
    O: int = 0
    L: int = 0
    E: int = 0
    scal: dict = field(default_factory=dict)      # SCALAR_REAL_KEYS + SOLVER_INT_KEYS
    safety: np.ndarray | None = None              # [N]
    safety_slack: np.ndarray | None = None        # [N]
    car: dict = field(default_factory=dict)       # CAR_VEC_KEYS -> [C]
    x0: np.ndarray | None = None                  # [C,6]  x,vx,ax,y,vy,ay
    ref: dict = field(default_factory=dict)       # CAR_STEP_KEYS -> [C,N]
    lim: dict = field(default_factory=dict)       # CAR_REGION_KEYS -> [C,R]
    initial_region: np.ndarray | None = None      # [C] 1-based
    possible_region: np.ndarray | None = None     # [C,R] int
    obs_edges: np.ndarray | None = None           # [O,N,L,4]
    obs_nedges: np.ndarray | None = None          # [O,N] int
    obs_soft: np.ndarray | None = None            # [O] int
    env_edges: np.ndarray | None = None           # [sum_e n_e, 4]
    env_off: np.ndarray | None = None             # [E+1] int
    frac: np.ndarray | None = None                # [R,4]
    poly: dict = field(default_factory=dict)      # POLY_KEYS -> [R,3]

    def copy(self) -> "FlatProblem":
        import copy
        return copy.deepcopy(self)


def _edges_from_set(s) -> np.ndarray:
    tuples = s["set"] if isinstance(s, dict) else s
    tuples = sorted(tuples, key=lambda t: t[0])
    return np.array([[t[1], t[2], t[3], t[4]] for t in tuples], dtype=np.float64).reshape(-1, 4)


def problem_from_dict(d: dict) -> FlatProblem:
    p = FlatProblem()
    p.N = int(d["NumSteps"])
    p.R = int(d["nr_regions"])
    p.C = int(d["NumCars"])
    p.O = int(d["nr_obstacles"])
    p.L = int(d["max_lines_obstacles"])
    p.E = int(d["nr_environments"])
    for k in SCALAR_REAL_KEYS:
        p.scal[k] = float(d[k])
    for k in SOLVER_INT_KEYS:
        p.scal[k] = int(d[k])
    p.safety = np.array(d["agent_safety_distance"], dtype=np.float64).reshape(p.N)
    p.safety_slack = np.array(d["agent_safety_distance_slack"], dtype=np.float64).reshape(p.N)
    for k in CAR_VEC_KEYS:
        p.car[k] = np.array(d[k], dtype=np.float64).reshape(p.C)
    p.x0 = np.array(d["IntitialState"], dtype=np.float64).reshape(p.C, 6)
    for k in CAR_STEP_KEYS:
        p.ref[k] = np.array(d[k], dtype=np.float64).reshape(p.C, p.N)
    for k in CAR_REGION_KEYS:
        p.lim[k] = np.array(d[k], dtype=np.float64).reshape(p.C, p.R)
    p.initial_region = np.array(d["initial_region"], dtype=np.int32).reshape(p.C)
    p.possible_region = np.array(d["possible_region"], dtype=np.int32).reshape(p.C, p.R)
    p.obs_edges = np.zeros((p.O, p.N, max(p.L, 1), 4))
    p.obs_nedges = np.zeros((p.O, p.N), dtype=np.int32)
    obs = d.get("ObstacleConvexPolygon", [])
    for o in range(p.O):
        for i in range(p.N):
            e = _edges_from_set(obs[o][i])
            if len(e) > p.L:
                raise ValueError("obstacle polygon with more edges than max_lines_obstacles")
            p.obs_nedges[o, i] = len(e)
            p.obs_edges[o, i, :len(e)] = e
    p.obs_edges = p.obs_edges[:, :, :p.L] if p.L > 0 else np.zeros((p.O, p.N, 0, 4))
    p.obs_soft = np.array(d.get("obstacle_is_soft", []), dtype=np.int32).reshape(p.O)
    env = d.get("MultiEnvironmentConvexPolygon", [])
    offs = [0]
    edges = []
    for e in range(p.E):
        ee = _edges_from_set(env[e])
        edges.append(ee)
        offs.append(offs[-1] + len(ee))
    p.env_edges = np.concatenate(edges, axis=0) if edges else np.zeros((0, 4))
    p.env_off = np.array(offs, dtype=np.int32)
    p.frac = np.array(d["fraction_parameters"], dtype=np.float64).reshape(p.R, 4)
    for k in POLY_KEYS:
        p.poly[k] = np.array(d[k], dtype=np.float64).reshape(p.R, 3)
    return p


def read_dat(path: str) -> FlatProblem:
    with open(path, "r") as f:
        return problem_from_dict(parse_dat_text(f.read()))


def _fmt(v: float, digits: int = 12) -> str:
    # the reference prints dumps with 12 display digits (src/cplex_wrapper.hpp:109-110)
    return repr(float(f"{v:.{digits}g}")) if v != int(v) or abs(v) > 1e15 else str(int(v))


def write_dat(p: FlatProblem, path: str, digits: int = 17) -> None:
    """OPL-printed dialect (whitespace separated), enough digits to round-trip doubles."""
    def num(v):
        return f"{float(v):.{digits}g}"

    def vec(a):
        return "[" + " ".join(num(x) for x in np.asarray(a).ravel()) + "]"

    def ivec(a):
        return "[" + " ".join(str(int(x)) for x in np.asarray(a).ravel()) + "]"

    def mat(a):
        return "[" + "\n".join(vec(r) for r in np.asarray(a)) + "]"

    def imat(a):
        return "[" + "\n".join(ivec(r) for r in np.asarray(a)) + "]"

    def edgeset(e):
        return "{" + "\n".join(
            f"<{k+1} {num(r[0])} {num(r[1])} {num(r[2])} {num(r[3])}>" for k, r in enumerate(e)) + "}"

    L = []
    L.append(f"NumSteps = {p.N};")
    L.append(f"nr_environments = {p.E};")
    L.append(f"nr_regions = {p.R};")
    L.append(f"nr_obstacles = {p.O};")
    L.append(f"max_lines_obstacles = {p.L};")
    L.append(f"NumCars = {p.C};")
    for k in ["max_solution_time", "relative_mip_gap_tolerance"]:
        L.append(f"{k} = {num(p.scal[k])};")
    for k in ["mipdisplay", "mipemphasis"]:
        L.append(f"{k} = {p.scal[k]};")
    L.append(f"relobjdif = {num(p.scal['relobjdif'])};")
    for k in ["cutpass", "probe", "repairtries", "rinsheur", "varsel", "mircuts", "parallelmode"]:
        L.append(f"{k} = {p.scal[k]};")
    for k in ["ts", "min_vel_x_y", "max_vel_x_y", "total_min_acc", "total_max_acc",
              "total_min_jerk", "total_max_jerk"]:
        L.append(f"{k} = {num(p.scal[k])};")
    L.append(f"agent_safety_distance = {vec(p.safety)};")
    L.append(f"agent_safety_distance_slack = {vec(p.safety_slack)};")
    L.append(f"maximum_slack = {num(p.scal['maximum_slack'])};")
    for k in CAR_VEC_KEYS[:8]:
        L.append(f"{k} = {vec(p.car[k])};")
    L.append(f"WEIGHTS_SLACK = {num(p.scal['WEIGHTS_SLACK'])};")
    L.append(f"WEIGHTS_SLACK_OBSTACLE = {num(p.scal['WEIGHTS_SLACK_OBSTACLE'])};")
    L.append(f"WheelBase = {vec(p.car['WheelBase'])};")
    L.append(f"CollisionRadius = {vec(p.car['CollisionRadius'])};")
    L.append(f"IntitialState = {mat(p.x0)};")
    for k in CAR_STEP_KEYS:
        L.append(f"{k} = {mat(p.ref[k])};")
    for k in CAR_REGION_KEYS:
        L.append(f"{k} = {mat(p.lim[k])};")
    L.append(f"initial_region = {ivec(p.initial_region)};")
    L.append(f"possible_region = {imat(p.possible_region)};")
    if p.O > 0:
        rows = []
        for o in range(p.O):
            rows.append("[" + " ".join(edgeset(p.obs_edges[o, i, :p.obs_nedges[o, i]])
                                       for i in range(p.N)) + "]")
        L.append("ObstacleConvexPolygon = [" + "\n".join(rows) + "];")
    else:
        L.append("ObstacleConvexPolygon = [];")
    L.append(f"obstacle_is_soft = {ivec(p.obs_soft)};")
    if p.E > 0:
        L.append("MultiEnvironmentConvexPolygon = [" + " ".join(
            edgeset(p.env_edges[p.env_off[e]:p.env_off[e + 1]]) for e in range(p.E)) + "];")
    else:
        L.append("MultiEnvironmentConvexPolygon = [];")
    L.append(f"fraction_parameters = {mat(p.frac)};")
    L.append(f"minimum_region_change_speed = {num(p.scal['minimum_region_change_speed'])};")
    for k in POLY_KEYS:
        L.append(f"{k} = {mat(p.poly[k])};")
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")
