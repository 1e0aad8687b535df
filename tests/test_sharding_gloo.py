"""Host logic of the multi-GPU path (scenario sharding) with world size 2 on the
gloo backend.  The per-rank solve is a stand-in (the oracle, which tests may call): what is
checked is the partition, the gather and the ordering -- the device solve itself is covered by
the -m gpu tests."""
import os
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import planner_miqp_b200  # noqa: F401
    from planner_miqp_b200.sharding import solve_sharded, shard_indices
    from planner_miqp_b200.scenarios import lane_following
    from oracle import oracle as O
    plans = [lane_following(seed=k, nr_steps=8).build() for k in range(5)]
    solved_here = []

    def fn(ps):
        out = [O.solve(p, gap_tol=1e-4, time_limit=30.0) for p in ps]
        solved_here.extend(range(len(ps)))
        return [o[0] for o in out], [o[1].objective for o in out]

    xs, objs = solve_sharded(plans, fn)
    assert len(solved_here) == len(shard_indices(5, rank, world))
    q.put((rank, [float(o) for o in objs], [float(x.sum()) for x in xs]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    sys.path.insert(0, ROOT)
    import planner_miqp_b200  # noqa: F401
    from planner_miqp_b200.sharding import shard_indices
    for count in (0, 1, 7, 4096):
        for world in (1, 2, 4, 8):
            parts = [shard_indices(count, r, world) for r in range(world)]
            flat = sorted(k for p in parts for k in p)
            assert flat == list(range(count))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_ranks_gloo_gather_in_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]      # same full result on both ranks
    assert len(res[0][1]) == 5 and all(o == o for o in res[0][1])
    # single-process reference
    sys.path.insert(0, ROOT)
    import planner_miqp_b200  # noqa: F401
    from planner_miqp_b200.scenarios import lane_following
    from oracle import oracle as O
    ref = [O.solve(lane_following(seed=k, nr_steps=8).build(), gap_tol=1e-4, time_limit=30.0)[1].objective for k in range(5)]
    assert res[0][1] == pytest.approx(ref, rel=1e-12)
