"""Frontier sharding on the GPU: the search in steps through the C ABI (miqp_b200_frontier_*) and the
driver planner-miqp_b200/sharding.py:solve_frontier_sharded.

* one process: the stepped search equals miqp_b200_batch_run;
* two processes on the gloo backend that share cuda:0 (the exchange goes through host buffers, so one GPU is enough):
  the sharded search returns the single-GPU optimum on every rank, bounds prove the gap, the winner's vector is feasible;
* two GPUs (skipped on a one-GPU box): the same over NCCL with the in-place min-all-reduce on the solver's device array."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario, parallel_lanes
from planner_miqp_b200.sharding import solve_frontier_sharded
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GAP = 1e-4


def plans_for_test():
    # single-car plans with obstacles (hundreds of nodes for the harder seeds) and one two-car plan with active collision rows
    return [obstacle_scenario(k).build() for k in (1, 7, 11)] + [parallel_lanes(2, 6, 4.8, stagger=1.0).build()]


def test_stepped_search_equals_batch_run():
    ps = plans_for_test()
    s = P.Solver()
    xs0, infos0 = s.solve_batch(ps, gap_tol=GAP, time_limit=120)
    st = {}
    xs1, infos1 = solve_frontier_sharded(s, ps, gap_tol=GAP, time_limit=120, stats=st)
    assert st["world"] == 1 and not st["split"]
    for a, b, xa, xb in zip(infos0, infos1, xs0, xs1):
        assert a.status == b.status == 0 and a.proven and b.proven
        assert a.objective == b.objective and a.nodes == b.nodes      # the same deterministic search
        assert np.array_equal(xa, xb)
    # the steps can also be driven by hand: a few rounds at a time
    s.upload(ps, gap_tol=GAP, time_limit=120)
    s.frontier_start()
    left, calls = len(ps), 0
    while left > 0:
        left = s.frontier_rounds(3)
        calls += 1
        assert calls < 1000
    s.frontier_finish()
    xs2, infos2 = s.fetch()
    assert [i.objective for i in infos2] == [i.objective for i in infos0]
    assert calls > 1
    s.close()


def _worker(rank, world, port, backend, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    ps = plans_for_test()
    s = P.Solver(device=dev)
    st = {}
    xs, infos = solve_frontier_sharded(s, ps, gap_tol=GAP, time_limit=120, ramp_rounds=5, exchange_every=3, stats=st)
    q.put((rank, [i.status for i in infos], [i.objective for i in infos], [i.best_bound for i in infos],
           [bool(i.proven) for i in infos], [x.copy() for x in xs], [i.nodes for i in infos], st))
    dist.barrier()
    dist.destroy_process_group()
    s.close()


def _run_two(backend):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ps = plans_for_test()
    s = P.Solver()
    xs0, infos0 = s.solve_batch(ps, gap_tol=GAP, time_limit=120)
    s.close()
    for r in res:
        assert r[7]["split"] and r[7]["world"] == 2 and r[7]["exchanges"] >= 1
        for k, p in enumerate(ps):
            assert r[1][k] == 0 and r[4][k], (k, r[1], r[4])
            # both searches prove 1e-4: the optima agree within the gap (usually to rounding)
            assert r[2][k] == pytest.approx(infos0[k].objective, rel=2e-4)
            assert r[3][k] <= r[2][k] + 1e-12 and r[2][k] - r[3][k] <= GAP * abs(r[2][k]) + 1e-12
            viol, worst = O.max_violation(p, r[5][k])
            assert viol <= 1e-6, (k, viol, worst)
            assert O.objective(p, r[5][k]) == pytest.approx(r[2][k], rel=1e-9, abs=1e-9)
    # identical combined results on both ranks
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]
    assert all(np.array_equal(a, b) for a, b in zip(res[0][5], res[1][5]))
    # both ranks worked on their own share
    assert res[0][7]["nodes_this_rank"] > 0 and res[1][7]["nodes_this_rank"] > 0


def test_two_ranks_one_gpu_gloo():
    _run_two("gloo")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_two_gpus_nccl():
    _run_two("nccl")
