"""Synthetic planning scenarios of the shapes named in BASELINE.json.

Every scenario is generated from a seed with the reference's default settings
(src/miqp_planner_data.hpp:190-242) through ``PlanBuilder``; no data set is involved.
"""
from __future__ import annotations

import math

import numpy as np

from .model_parameters import PlanBuilder, Settings, default_settings


def settings_for(nr_regions=32, nr_steps=40, ts=0.25, gap=1e-4, time_limit=10.0) -> Settings:
    s = default_settings()
    s.nr_regions, s.nr_steps, s.ts = nr_regions, nr_steps, ts
    s.relative_mip_gap_tolerance, s.max_solution_time = gap, time_limit
    return s


def lane_following(seed: int = 0, nr_regions=16, nr_steps=20):
    """config 1c: miqp_planner_test `plan1` shape (test/miqp_planner_test.cc:279-308)."""
    rng = np.random.default_rng(seed)
    s = settings_for(nr_regions, nr_steps)
    b = PlanBuilder(s)
    v0 = 4.0 + (rng.uniform(-1, 1) if seed else 0.0)
    b.add_car([0, v0, 0, rng.uniform(-0.5, 0.5) if seed else 0.0, 0.1, 0], [[0, 0], [200, 0]], 5.0, 1.0)
    b.add_environment_polygon([[-49, -49], [249, -49], [249, 49], [-49, 49]])
    return b


def obstacle_scenario(seed: int = 0, nr_regions=32, nr_steps=40, ts=0.25, n_static=1, n_dynamic=1, soft=False):
    """config 2: single agent, static + dynamic-occupancy obstacles, N=40, 32 fitted regions.

    Environment rectangle [-10,200]x[-10,20] shrunk by the collision radius; the ego starts at
    x0=[0,5,0,0,0.1,0] and follows the x axis at 5 m/s; static 1x1 boxes at (20+U[0,20], +-1.5)
    (centre offset +-U[0.3,1.5] so that the inflated box blocks the lane) and dynamic 1x1 boxes
    from (10, +-U[2.5,4]) moving along +x at U[2,6] m/s, all inflated by the collision radius
    to 3x3 squares with 4 edges (src/miqp_planner.cpp:444-488)."""
    rng = np.random.default_rng(1000 + seed)
    s = settings_for(nr_regions, nr_steps, ts)
    b = PlanBuilder(s)
    v0 = 5.0 + rng.uniform(-1.0, 1.0)
    b.add_car([0, v0, 0, rng.uniform(-0.3, 0.3), 0.1, 0], [[0, 0], [400, 0]], 5.0, 1.0)
    r = s.collisionRadius
    b.add_environment_polygon([[-10 + r, -10 + r], [200 - r, -10 + r], [200 - r, 20 - r], [-10 + r, 20 - r]])
    for k in range(n_static):
        side = 1.0 if rng.uniform() < 0.5 else -1.0
        b.add_box_obstacle([[20.0 + rng.uniform(0, 20) + 25.0 * k, side * rng.uniform(0.3, 1.5), 0.0]], 1.0, 1.0, soft=soft)
    for k in range(n_dynamic):
        v = rng.uniform(2.0, 6.0)
        y = (1.0 if rng.uniform() < 0.5 else -1.0) * rng.uniform(2.5, 4.0)
        centers = [[10.0 + 15.0 * k + v * s.ts * i, y, 0.0] for i in range(nr_steps)]
        b.add_box_obstacle(centers, 1.0, 1.0, soft=soft)
    return b


def random_single_agent(seed: int, nr_regions=16, nr_steps=20):
    """config 4 restricted to one agent: straight or single-bend reference, 0-2 obstacles."""
    rng = np.random.default_rng(5000 + seed)
    s = settings_for(nr_regions, nr_steps)
    b = PlanBuilder(s)
    bend = math.radians(rng.uniform(-30, 30)) if rng.uniform() < 0.5 else 0.0
    ref = [[0, 0], [30, 0], [30 + 170 * math.cos(bend), 170 * math.sin(bend)]]
    v0 = rng.uniform(3, 8)
    b.add_car([0, v0, 0, rng.uniform(-0.5, 0.5), 0.05, 0], ref, v0, 1.0)
    if rng.uniform() < 0.5:
        b.add_environment_polygon([[-20, -60], [220, -60], [220, 60], [-20, 60]])
    for k in range(int(rng.integers(0, 3))):
        side = 1.0 if rng.uniform() < 0.5 else -1.0
        b.add_box_obstacle([[15.0 + rng.uniform(0, 20) + 20 * k, side * rng.uniform(1.0, 2.0), 0.0]], 1.0, 1.0)
    return b


def parallel_lanes(n_cars: int, nr_steps=5, lane_offset=4.5, nr_regions=16, stagger=0.0, v=5.0):
    """n cars side by side on parallel straight references (shape of the reference's n_agents test,
    test/miqp_planner_test.cc:1528-1565: N=5).  With lane_offset below RR + slack (= 5 m) the
    collision rows and their slacks are active."""
    s = settings_for(nr_regions, nr_steps)
    b = PlanBuilder(s)
    for c in range(n_cars):
        y = lane_offset * c
        b.add_car([stagger * c, v, 0, y, 0.0, 0], [[-20, y], [100, y]], v, 1.0)
    return b


def two_agent_merge(seed: int = 0, nr_regions=16, nr_steps=20):
    """config 3: two-agent cooperative merge (joint MIQP, lambda = 0.5): the ego follows the x axis,
    the other car comes from an ending lane at y = -3.5 that merges into it."""
    rng = np.random.default_rng(3000 + seed)
    s = settings_for(nr_regions, nr_steps)
    b = PlanBuilder(s)
    gap = rng.uniform(-5.0, 5.0)
    b.add_car([0, rng.uniform(4, 6), 0, 0, 0.0, 0], [[0, 0], [100, 0]], 5.0, 1.0)
    b.add_car([gap, rng.uniform(4, 6), 0, -3.5, 0.0, 0], [[-20, -3.5], [30, -3.5], [50, 0], [100, 0]], 5.0, 1.0)
    r = s.collisionRadius
    b.add_environment_polygon([[-30 + r, -8 + r], [130 - r, -8 + r], [130 - r, 8 - r], [-30 + r, 8 - r]])
    return b


def random_scenario(seed: int, nr_regions=16, nr_steps=20, max_agents=4):
    """config 4: 1-4 agents on parallel lanes 6 m apart (collision rows present, rarely active),
    straight or single-bend references, 0-2 static obstacles in the ego lane."""
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(1, max_agents + 1))
    if n == 1:
        return random_single_agent(seed, nr_regions, nr_steps)
    s = settings_for(nr_regions, nr_steps)
    b = PlanBuilder(s)
    bend = math.radians(rng.uniform(-20, 20)) if rng.uniform() < 0.5 else 0.0
    for c in range(n):
        y0 = 6.0 * c
        ref = [[-10, y0], [30, y0], [30 + 170 * math.cos(bend), y0 + 170 * math.sin(bend)]]
        v0 = rng.uniform(3, 8)
        b.add_car([rng.uniform(-3, 3), v0, 0, y0 + rng.uniform(-0.3, 0.3), 0.05, 0], ref, v0, 1.0)
    for k in range(int(rng.integers(0, 2))):
        side = 1.0 if rng.uniform() < 0.5 else -1.0
        b.add_box_obstacle([[20.0 + rng.uniform(0, 15), side * rng.uniform(1.2, 2.0), 0.0]], 1.0, 1.0)
    return b


def intersection(seed: int = 0, n_cars: int = 8, nr_steps=40, nr_regions=64, ts=0.25, time_limit=10.0, gap=1e-4):
    """config 5: n-agent four-way intersection (joint MIQP over all cars, lambda = 0.5), N=40, 64 fitted regions
    (fitting table 64/10/1: max_velocity_fitting 10 m/s, minimum_region_change_speed 1 m/s).

    Two lanes per road (3.5 m wide, right-hand traffic), cars approach from the four arms -- car c comes from arm c mod 4, the
    second car of an arm follows 12-16 m behind the first -- 14-24 m before the conflict area at 4-6 m/s, every reference goes
    straight across.  The crossing order of every conflicting pair is what the collision-side binaries of
    agent_collision_constraints.mod:38-73 decide; the frontier of that search is what is sharded over the GPUs."""
    rng = np.random.default_rng(9000 + seed)
    s = settings_for(nr_regions, nr_steps, ts, gap, time_limit)
    s.max_velocity_fitting, s.minimum_region_change_speed = 10.0, 1.0
    b = PlanBuilder(s)
    lane = 1.75
    arms = [((-1.0, 0.0), (0.0, -lane)), ((0.0, -1.0), (lane, 0.0)), ((1.0, 0.0), (0.0, lane)), ((0.0, 1.0), (-lane, 0.0))]
    for c in range(n_cars):
        (ax, ay), (ox, oy) = arms[c % 4]                  # arm direction (from the centre outwards), lane offset
        d = rng.uniform(14.0, 24.0) + (c // 4) * rng.uniform(12.0, 16.0)
        v = rng.uniform(4.0, 6.0)
        px, py = ax * d + ox, ay * d + oy
        hx, hy = -ax, -ay                                  # heading: towards the centre
        ref = [[px - hx * 5.0, py - hy * 5.0], [px + hx * 160.0, py + hy * 160.0]]
        b.add_car([px, v * hx, 0.0, py, v * hy, 0.0], ref, 5.0, 1.0)
    r = s.collisionRadius
    b.add_environment_polygon([[-120 + r, -120 + r], [120 - r, -120 + r], [120 - r, 120 - r], [-120 + r, 120 - r]])
    return b


def advance_obstacle_scenario(builder: PlanBuilder, plan, x) -> PlanBuilder:
    """Next planning cycle of a config-2 scenario (receding horizon): the car state becomes step 1 of
    the solution x, every obstacle prediction moves one step ahead (the last pose is extrapolated
    with the last displacement), reference and environment stay."""
    from .results import block_views
    v = block_views(plan, x)
    nb = PlanBuilder(builder.s)
    for c, car in enumerate(builder.cars):
        st = [v["pos_x"][c, 1], v["vel_x"][c, 1], v["acc_x"][c, 1], v["pos_y"][c, 1], v["vel_y"][c, 1], v["acc_y"][c, 1]]
        nb.add_car(st, car["ref"], car["v_des"], car["ds"], car["track"])
    nb.envs = [e.copy() for e in builder.envs]
    for polys, soft in builder.obstacles:
        last = polys[-1] + (polys[-1] - polys[-2]) if len(polys) > 1 else polys[-1]
        nb.obstacles.append((list(polys[1:]) + [last], soft))
    return nb
