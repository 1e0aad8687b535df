"""Scenario sharding across the GPUs of one box.

Plans are independent units: plan s of a batch goes to rank s mod world.  There is no collective
on the data path; the per-plan results (status, objective, gap, solution vector) are gathered on
every rank at the end with one all_gather_object over the process group (NCCL ranks use a gloo
side group for Python objects; on CPU test boxes the group is gloo).  One process per GPU,
launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment).
"""
from __future__ import annotations

import os
from typing import Callable, Sequence


def shard_indices(count: int, rank: int, world: int) -> list[int]:
    """indices of the plans rank `rank` solves: s mod world == rank (round robin balances the
    mix of cheap and expensive scenarios better than contiguous blocks)"""
    return list(range(rank, count, world))


def default_solve_fn(gap_tol=None, time_limit=None):
    """solves a shard on this rank's GPU (LOCAL_RANK) through the C ABI; no CPU fallback"""
    from .capi import Solver
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    solver = Solver(device=dev)

    def fn(plans):
        if not plans:
            return [], []
        return solver.solve_batch(plans, gap_tol=gap_tol, time_limit=time_limit)
    return fn


def solve_sharded(problems: Sequence, solve_fn: Callable, group=None):
    """Every rank passes the same `problems`; returns (xs, infos) for ALL plans on every rank,
    in the order of `problems`."""
    import torch.distributed as dist
    if not dist.is_initialized():
        return solve_fn(list(problems))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_indices(len(problems), rank, world)
    xs, infos = solve_fn([problems[k] for k in mine])
    payload = [(k, x, i) for k, x, i in zip(mine, xs, infos)]
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    out_x, out_i = [None] * len(problems), [None] * len(problems)
    for part in gathered:
        for k, x, i in part:
            out_x[k], out_i[k] = x, i
    return out_x, out_i
