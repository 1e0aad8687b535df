"""Host-side preparation of a plan: Settings -> flat ModelParameters.

Mirrors, without BARK/boost/Eigen, the part of the reference that turns a few user inputs
into the arrays the MIQP consumes:

* ``Settings`` / ``default_settings``: ``MiqpPlannerSettings`` and ``DefaultSettings()``
  (src/miqp_planner_settings.h:28-78, src/miqp_planner_data.hpp:190-242).
* ``ParameterPreparer``: fraction parameters, mean angles and rotated per-region acc/jerk
  boxes (common/parameter/parameter_preparer.cpp:37-143), fitted polynomial tables
  (data/fitting_tables.json, extracted from fitting_polynomial_parameters.hpp).
* region helpers (common/parameter/regions.cpp:16-127).
* ``reference_trajectory``: the speed-ramp walk along a poly-line centre line
  (common/reference/reference_trajectory_generator.cpp:51-148; no curvature-dependent speed,
  as in the planner, src/miqp_planner.cpp:252,258).
* ``PlanBuilder``: ``MiqpPlanner``'s parameter plumbing: constructor (src/miqp_planner.cpp:62-116),
  AddCar/UpdateCar (:180-390), AddObstacle/CreateMiqpObstacle for box shapes (:405-488),
  environment polygons as CCW vertex lists (common/geometry/geometry.cpp:126-139), initial
  region (:635-646, :696-712).

This is input preparation (a few kB per plan, CPU); the MIQP itself runs on the GPU.
"""
from __future__ import annotations

import copy
import json
import math
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32


@dataclass
class Settings:
    nr_regions: int = 16
    nr_steps: int = 20
    nr_neighbouring_possible_regions: int = 1
    ts: float = 0.25
    precision: int = 12
    constant_agent_safety_distance_slack: float = 3.0
    minimum_region_change_speed: float = 2.0
    lambda_: float = 0.5
    wheelBase: float = 2.8
    collisionRadius: float = 1.0
    slackWeight: float = 30.0
    slackWeightObstacle: float = 2000.0
    jerkWeight: float = 1.0
    positionWeight: float = 2.0
    velocityWeight: float = 0.0
    acclerationWeight: float = 0.0          # (sic) the misspelling is API
    accLonMaxLimit: float = 2.0
    accLonMinLimit: float = -4.0
    jerkLonMaxLimit: float = 3.0
    accLatMinMaxLimit: float = 1.6
    jerkLatMinMaxLimit: float = 1.4
    simplificationDistanceMap: float = 0.2
    simplificationDistanceReferenceLine: float = 0.05
    bufferReference: float = 1.0
    buffer_for_merging_tolerance: float = 0.1
    refLineInterpInc: float = 0.2
    additionalStepsForReferenceLongerHorizon: int = 4
    max_solution_time: float = 10.0
    relative_mip_gap_tolerance: float = 0.1
    mipdisplay: int = 2
    mipemphasis: int = 0
    relobjdif: float = 0.0
    cutpass: int = 0
    probe: int = 0
    repairtries: int = 0
    rinsheur: int = 0
    varsel: int = 0
    mircuts: int = 0
    useSos: bool = False
    useBranchingPriorities: bool = False
    warmstartType: int = 0
    parallelMode: int = 1
    max_velocity_fitting: float = 20.0


def default_settings() -> Settings:
    return Settings()


@dataclass
class FlatPlan:
    """Flat ModelParameters (same attribute names the C-ABI binding reads)."""
    N: int = 0
    R: int = 0
    C: int = 0
    O: int = 0
    L: int = 0
    E: int = 0
    scal: dict = field(default_factory=dict)
    safety: np.ndarray | None = None
    safety_slack: np.ndarray | None = None
    car: dict = field(default_factory=dict)
    x0: np.ndarray | None = None
    ref: dict = field(default_factory=dict)
    lim: dict = field(default_factory=dict)
    initial_region: np.ndarray | None = None
    possible_region: np.ndarray | None = None
    obs_edges: np.ndarray | None = None
    obs_nedges: np.ndarray | None = None
    obs_soft: np.ndarray | None = None
    env_edges: np.ndarray | None = None
    env_off: np.ndarray | None = None
    frac: np.ndarray | None = None
    poly: dict = field(default_factory=dict)

    def copy(self) -> "FlatPlan":
        return copy.deepcopy(self)


_TABLES = None


def fitting_tables(nr_regions: int, vmax: float, vmin: float) -> dict:
    """POLY_* tables [R][3]; raises ValueError for combinations the reference does not ship
    (fitting_polynomial_parameters.hpp:33-44)."""
    global _TABLES
    if _TABLES is None:
        with open(os.path.join(_HERE, "data", "fitting_tables.json")) as f:
            _TABLES = json.load(f)["tables"]
    key = f"{int(nr_regions)},{int(vmax)},{int(vmin)}"
    if key not in _TABLES or float(vmax) != int(vmax) or float(vmin) != int(vmin):
        raise ValueError("Invalid number of regions or velocity!")
    return {k: np.array(v, dtype=np.float64) for k, v in _TABLES[key].items()}


def wrap_to_2pi(a: float) -> float:
    a = math.fmod(a, 2.0 * math.pi)
    if a < 0:
        a += 2.0 * math.pi
    return a


class ParameterPreparer:
    def __init__(self, s: Settings):
        self.R = s.nr_regions
        self.vmax = float(_f32(s.max_velocity_fitting))
        self.poly = fitting_tables(s.nr_regions, s.max_velocity_fitting, s.minimum_region_change_speed)
        # straight-driving limits are floats in the reference (vehicle_parameters.hpp:70-88)
        self.acc = (float(_f32(s.accLonMinLimit)), float(_f32(s.accLonMaxLimit)),
                    -float(_f32(s.accLatMinMaxLimit)), float(_f32(s.accLatMinMaxLimit)))
        self.jerk = (-float(_f32(s.jerkLonMaxLimit)), float(_f32(s.jerkLonMaxLimit)),
                     -float(_f32(s.jerkLatMinMaxLimit)), float(_f32(s.jerkLatMinMaxLimit)))
        alpha = np.linspace(0.0, 2.0 * math.pi, self.R + 1)
        col = np.stack([self.vmax * np.cos(alpha[:-1]), self.vmax * np.sin(alpha[:-1])], axis=1)
        self.frac = np.zeros((self.R, 4))
        self.frac[:, :2] = col
        self.frac[:-1, 2:] = col[1:]
        self.frac[-1, 2:] = col[0]
        self.mean_angles = []
        for j in range(self.R):
            a1 = wrap_to_2pi(math.atan2(self.frac[j, 1], self.frac[j, 0]))
            a2 = wrap_to_2pi(math.atan2(self.frac[j, 3], self.frac[j, 2]))
            if j + 1 == self.R:
                a2 += 2.0 * math.pi
            self.mean_angles.append((a1 + a2) / 2.0)

    @staticmethod
    def _rotate_limits(lim, angle):
        """RotateLimitVectors: extreme x/y over the four rotated corners, in float arithmetic."""
        lon_min, lon_max, lat_min, lat_max = lim
        th = float(_f32(angle))
        c, s = math.cos(th), math.sin(th)
        xs, ys = [], []
        for lx in (lon_max, lon_min):
            for ly in (lat_max, lat_min):
                xs.append(float(_f32(lx * c - ly * s)))
                ys.append(float(_f32(lx * s + ly * c)))
        return min(xs), max(xs), min(ys), max(ys)

    def limits_per_region(self, lim):
        out = np.zeros((4, self.R))      # min_x, max_x, min_y, max_y
        for j, a in enumerate(self.mean_angles):
            out[:, j] = self._rotate_limits(lim, a)
        return out


def region_indices(frac: np.ndarray, vx: float, vy: float) -> list[int]:
    """CalculateRegionIdx: all regions whose wedge contains (vx, vy) up to eps=1e-3."""
    eps = float(_f32(1e-3))
    vx, vy = float(_f32(vx)), float(_f32(vy))
    out = []
    for j in range(frac.shape[0]):
        below_ub = frac[j, 2] * vy <= frac[j, 3] * vx + eps
        above_lb = frac[j, 0] * vy >= frac[j, 1] * vx - eps
        if below_ub and above_lb:
            out.append(j)
    return out


def reserve_neighbor_regions(mask: np.ndarray, expansions: int) -> bool:
    """ReserveNeighborRegions (regions.cpp:75-114), in place on a 0/1 vector."""
    for _ in range(max(expansions, 0)):
        first = last = None
        s = len(mask)
        for i in range(s):
            if mask[i] == 1:
                if i >= 1 and mask[i - 1] == 0:
                    first = i - 1
                if i + 1 < s and mask[i + 1] == 0:
                    last = i + 1
                if i == 0 and mask[s - 1] == 0:
                    first = s - 1
                if i == s - 1 and mask[0] == 0:
                    last = 0
        if first is None or last is None:
            return False
        mask[first] = 1
        mask[last] = 1
    return True


class PolyLine:
    """Centre line with arclength (the subset of bark::geometry::Line the generator needs)."""

    def __init__(self, pts, interp_inc: float = 0.0):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        if interp_inc > 0:
            seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
            s = np.concatenate([[0.0], np.cumsum(seg)])
            n = max(int(math.ceil(s[-1] / interp_inc)), 1)
            ss = np.linspace(0.0, s[-1], n + 1)
            pts = np.stack([np.interp(ss, s, pts[:, 0]), np.interp(ss, s, pts[:, 1])], axis=1)
        self.pts = pts
        seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
        self.s = np.concatenate([[0.0], np.cumsum(seg)])

    def nearest_s(self, x, y):
        a, b = self.pts[:-1], self.pts[1:]
        d = b - a
        L2 = np.einsum("ij,ij->i", d, d)
        t = np.where(L2 > 0, ((x - a[:, 0]) * d[:, 0] + (y - a[:, 1]) * d[:, 1]) / np.where(L2 > 0, L2, 1.0), 0.0)
        t = np.clip(t, 0.0, 1.0)
        q = a + t[:, None] * d
        dist = (q[:, 0] - x) ** 2 + (q[:, 1] - y) ** 2
        k = int(np.argmin(dist))          # first minimum, as a sequential scan with '<' finds it
        return float(self.s[k] + t[k] * math.sqrt(L2[k]))

    def point_at(self, s):
        s = min(max(s, 0.0), self.s[-1])
        return float(np.interp(s, self.s, self.pts[:, 0])), float(np.interp(s, self.s, self.pts[:, 1]))

    def angle_at(self, s):
        s = min(max(s, 0.0), self.s[-1])
        k = int(np.searchsorted(self.s, s, side="right") - 1)
        k = min(max(k, 0), len(self.pts) - 2)
        d = self.pts[k + 1] - self.pts[k]
        return math.atan2(d[1], d[0])


def reference_trajectory(line: PolyLine, x, y, v0, n_points, dt, vel_desired, delta_s_desired):
    """Rows (x, y, theta, v); row 0 is the current state (theta = heading handed in by the caller)."""
    s_start = line.nearest_s(x, y)
    s_end = line.s[-1]
    s_des = min(s_end, s_start + delta_s_desired)
    vel_0 = vel_desired if delta_s_desired <= 0.0 else v0
    vel_i, vel_end = vel_0, vel_desired
    if vel_i * n_points * dt + s_start > s_end:
        if vel_end > 0 or s_des > s_end:
            vel_end = 0.0
            s_des = s_end
    out = np.zeros((n_points, 4))
    s_i = s_start
    for i in range(1, n_points):
        s_i += vel_i * dt
        px, py = line.point_at(s_i)
        th = line.angle_at(s_i)
        if (s_des - s_start) < 1e-2:
            vel_i = 0.0
        elif s_i > s_des:
            vel_i = vel_end
        elif s_i < s_start:
            vel_i = vel_0
        else:
            vel_i = (vel_end - vel_0) / (s_des - s_start) * (s_i - s_start) + vel_0
        out[i] = (px, py, th, vel_i)
    return out


def box_polygon_ccw(cx, cy, theta, length, width, inflate):
    """Inflated bounding box of a shape at a pose (CreateMiqpObstacle keeps 4 edges), CCW,
    without the repeated closing vertex (Polygon2MiqpPolygonDefinition)."""
    hl, hw = length / 2.0 + inflate, width / 2.0 + inflate
    corners = np.array([[-hl, -hw], [hl, -hw], [hl, hw], [-hl, hw]])
    c, s = math.cos(theta), math.sin(theta)
    rot = np.array([[c, -s], [s, c]])
    return corners @ rot.T + np.array([cx, cy])


def edges_from_vertices(v):
    """addLineSet: edge k runs from vertex k to vertex k+1, the last edge closes the polygon
    (src/model_input_data_source.cpp:167-178)."""
    v = np.asarray(v, dtype=np.float64).reshape(-1, 2)
    nxt = np.roll(v, -1, axis=0)
    return np.concatenate([v, nxt], axis=1)


def round_reals(a, precision: int):
    """RoundWithPrecision (src/model_input_data_source.hpp:74-79) with precision-2 decimals."""
    scale = 10.0 ** (precision - 2)
    return np.round(np.asarray(a, dtype=np.float64) * scale) / scale


class PlanBuilder:
    """MiqpPlanner's parameter plumbing for straight / poly-line references and box obstacles."""

    EPS = 1e-6

    def __init__(self, settings: Settings | None = None):
        self.s = settings or default_settings()
        self.prep = ParameterPreparer(self.s)
        self.cars = []        # dicts
        self.obstacles = []   # (list of [k,2] vertex arrays per step, soft)
        self.envs = []        # [k,2] CCW vertex arrays

    def add_car(self, state6, ref_points, v_des, delta_s_des=1.0, track_reference_positions=True):
        self.cars.append(dict(state=np.asarray(state6, dtype=np.float64), ref=np.asarray(ref_points, dtype=np.float64),
                              v_des=float(v_des), ds=float(delta_s_des), track=bool(track_reference_positions)))
        return len(self.cars) - 1

    def update_car(self, idx, state6):
        self.cars[idx]["state"] = np.asarray(state6, dtype=np.float64)

    def add_environment_polygon(self, vertices_ccw):
        self.envs.append(np.asarray(vertices_ccw, dtype=np.float64).reshape(-1, 2))

    def add_box_obstacle(self, centers_xytheta, length, width, soft=False):
        """centers_xytheta: [N,3] predicted poses; static obstacles repeat one pose."""
        c = np.asarray(centers_xytheta, dtype=np.float64).reshape(-1, 3)
        if c.shape[0] == 1:
            c = np.repeat(c, self.s.nr_steps, axis=0)
        polys = [box_polygon_ccw(x, y, th, length, width, self.s.collisionRadius) for x, y, th in c]
        self.obstacles.append((polys, bool(soft)))
        return len(self.obstacles) - 1

    def build(self, initial_region_choice: int = 0) -> FlatPlan:
        s, prep = self.s, self.prep
        N, R, Cn = s.nr_steps, s.nr_regions, len(self.cars)
        p = FlatPlan(N=N, R=R, C=Cn, O=len(self.obstacles), L=4 if self.obstacles else 0, E=len(self.envs))
        f = lambda v: float(_f32(v))     # noqa: E731  (ModelParameters scalars are float)
        acc = prep.limits_per_region(prep.acc)
        jerk = prep.limits_per_region(prep.jerk)
        p.lim = {"min_acc_x": np.tile(acc[0], (Cn, 1)), "max_acc_x": np.tile(acc[1], (Cn, 1)),
                 "min_acc_y": np.tile(acc[2], (Cn, 1)), "max_acc_y": np.tile(acc[3], (Cn, 1)),
                 "min_jerk_x": np.tile(jerk[0], (Cn, 1)), "max_jerk_x": np.tile(jerk[1], (Cn, 1)),
                 "min_jerk_y": np.tile(jerk[2], (Cn, 1)), "max_jerk_y": np.tile(jerk[3], (Cn, 1))}
        p.scal = {
            "ts": f(s.ts), "max_solution_time": f(s.max_solution_time),
            "relative_mip_gap_tolerance": f(s.relative_mip_gap_tolerance), "relobjdif": f(s.relobjdif),
            "min_vel_x_y": f(-(s.max_velocity_fitting + self.EPS)), "max_vel_x_y": f(s.max_velocity_fitting + self.EPS),
            "total_max_acc": f(max(acc[1].max(), acc[3].max()) + self.EPS),
            "total_min_acc": f(min(acc[0].min(), acc[2].min()) - self.EPS),
            "total_max_jerk": f(max(jerk[1].max(), jerk[3].max()) + self.EPS),
            "total_min_jerk": f(min(jerk[0].min(), jerk[2].min()) - self.EPS),
            "maximum_slack": f(s.constant_agent_safety_distance_slack),
            "WEIGHTS_SLACK": f(s.slackWeight), "WEIGHTS_SLACK_OBSTACLE": f(s.slackWeightObstacle),
            "minimum_region_change_speed": f(s.minimum_region_change_speed),
            "mipdisplay": s.mipdisplay, "mipemphasis": s.mipemphasis, "cutpass": s.cutpass, "probe": s.probe,
            "repairtries": s.repairtries, "rinsheur": s.rinsheur, "varsel": s.varsel, "mircuts": s.mircuts,
            "parallelmode": s.parallelMode,
        }
        p.safety = np.zeros(N)
        p.safety_slack = np.full(N, f(s.constant_agent_safety_distance_slack))
        p.frac = prep.frac.copy()
        p.poly = {k: v.copy() for k, v in prep.poly.items()}
        p.x0 = np.zeros((Cn, 6))
        p.ref = {k: np.zeros((Cn, N)) for k in ("x_ref", "vx_ref", "y_ref", "vy_ref")}
        p.possible_region = np.zeros((Cn, R), dtype=np.int32)
        p.initial_region = np.zeros(Cn, dtype=np.int32)
        names = ["WEIGHTS_POS_X", "WEIGHTS_VEL_X", "WEIGHTS_ACC_X", "WEIGHTS_POS_Y", "WEIGHTS_VEL_Y",
                 "WEIGHTS_ACC_Y", "WEIGHTS_JERK_X", "WEIGHTS_JERK_Y"]
        p.car = {k: np.zeros(Cn) for k in names}
        p.car["WheelBase"] = np.full(Cn, f(s.wheelBase))
        p.car["CollisionRadius"] = np.full(Cn, f(s.collisionRadius))
        combos = []
        for c, car in enumerate(self.cars):
            st = car["state"]
            p.x0[c] = st
            line = PolyLine(car["ref"], s.refLineInterpInc)
            v0 = math.hypot(st[1], st[4])
            traj = reference_trajectory(line, st[0], st[3], v0, N, s.ts, car["v_des"], car["ds"])
            traj[0] = (st[0], st[3], math.atan2(st[4], st[1]), v0)
            p.ref["x_ref"][c] = traj[:, 0]
            p.ref["y_ref"][c] = traj[:, 1]
            p.ref["vx_ref"][c] = traj[:, 3] * np.cos(traj[:, 2])
            p.ref["vy_ref"][c] = traj[:, 3] * np.sin(traj[:, 2])
            longer = reference_trajectory(line, st[0], st[3], v0, N + s.additionalStepsForReferenceLongerHorizon,
                                          s.ts, car["v_des"], car["ds"])
            longer[0, 2] = math.atan2(st[4], st[1])
            mask = np.zeros(R, dtype=np.int32)
            for th in longer[:, 2]:
                for j in region_indices(prep.frac, math.cos(th), math.sin(th)):
                    mask[j] = 1
            reserve_neighbor_regions(mask, s.nr_neighbouring_possible_regions)
            scale = s.lambda_ if c == 0 else (1.0 - s.lambda_) / max(Cn - 1, 1)
            if car["track"]:
                wp, wv = scale * s.positionWeight, scale * s.velocityWeight
            else:
                wp, wv = 0.0, 2.0
            for k, v in zip(names, (wp, wv, scale * s.acclerationWeight, wp, wv, scale * s.acclerationWeight,
                                    scale * s.jerkWeight, scale * s.jerkWeight)):
                p.car[k][c] = f(v)
            regs = region_indices(prep.frac, st[1], st[4])
            combos.append(regs)
            j0 = regs[min(initial_region_choice, len(regs) - 1)]
            p.initial_region[c] = j0 + 1           # 1-based for OPL (src/miqp_planner.cpp:696-700)
            mask[j0] = 1                            # start region forced possible (:703-712)
            p.possible_region[c] = mask
        self.initial_region_combinations = combos
        O, L = p.O, max(p.L, 1)
        p.obs_edges = np.zeros((O, N, p.L, 4)) if O else np.zeros((0, N, 0, 4))
        p.obs_nedges = np.zeros((O, N), dtype=np.int32)
        p.obs_soft = np.zeros(O, dtype=np.int32)
        for o, (polys, soft) in enumerate(self.obstacles):
            p.obs_soft[o] = int(soft)
            for i in range(N):
                e = edges_from_vertices(polys[min(i, len(polys) - 1)])
                p.obs_nedges[o, i] = len(e)
                p.obs_edges[o, i, :len(e)] = e
        offs, edges = [0], []
        for v in self.envs:
            e = edges_from_vertices(v)
            edges.append(e)
            offs.append(offs[-1] + len(e))
        p.env_edges = np.concatenate(edges, axis=0) if edges else np.zeros((0, 4))
        p.env_off = np.array(offs, dtype=np.int32)
        # every real enters the model rounded to precision-2 decimals
        pr = s.precision
        for k in list(p.scal):
            if isinstance(p.scal[k], float):
                p.scal[k] = float(round_reals(p.scal[k], pr))
        for d in (p.car, p.ref, p.lim, p.poly):
            for k in d:
                d[k] = round_reals(d[k], pr)
        p.safety, p.safety_slack = round_reals(p.safety, pr), round_reals(p.safety_slack, pr)
        p.x0, p.frac = round_reals(p.x0, pr), round_reals(p.frac, pr)
        p.obs_edges, p.env_edges = round_reals(p.obs_edges, pr), round_reals(p.env_edges, pr)
        return p
