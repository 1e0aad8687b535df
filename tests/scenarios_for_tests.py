"""Problem variants derived from the reference fixtures, for parity tests only."""
import copy

import numpy as np


def multi_car_variant(p, cars=2, soft=True, two_env=True):
    """Replicates the single car of a fixture into `cars` laterally shifted cars, adds a second
    environment polygon and makes the obstacle soft: exercises the agent_collision rows, the
    multi-environment sums and the soft-obstacle slack columns of the row generator."""
    q = copy.deepcopy(p)
    C = cars
    q.C = C
    for k in list(q.car):
        q.car[k] = np.repeat(q.car[k][:1], C)
    q.x0 = np.repeat(q.x0[:1], C, axis=0).copy()
    for c in range(C):
        q.x0[c, 3] += 4.0 * c
    for k in list(q.ref):
        q.ref[k] = np.repeat(q.ref[k][:1], C, axis=0).copy()
    for c in range(C):
        q.ref["y_ref"][c] += 4.0 * c
    for k in list(q.lim):
        q.lim[k] = np.repeat(q.lim[k][:1], C, axis=0)
    q.initial_region = np.repeat(q.initial_region[:1], C)
    q.possible_region = np.repeat(q.possible_region[:1], C, axis=0).copy()
    if C > 1:
        q.possible_region[1, :] = 0
        q.possible_region[1, :5] = 1
    if soft and q.O > 0:
        q.obs_soft = np.ones_like(q.obs_soft)
    if two_env and q.E > 0:
        e = q.env_edges[q.env_off[0]:q.env_off[1]].copy()
        e2 = e.copy()
        e2[:, [0, 2]] += 30.0
        tri = np.array([[0, 0, 5, 0], [5, 0, 0, 5], [0, 5, 0, 0]], dtype=float)
        q.env_edges = np.concatenate([e, e2, tri], axis=0)
        q.env_off = np.array([0, len(e), 2 * len(e), 2 * len(e) + 3], dtype=np.int32)
        q.E = 3
    q.safety = np.linspace(0.0, 0.5, q.N)
    return q
