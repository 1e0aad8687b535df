"""Host logic of frontier sharding (planner-miqp_b200/sharding.py:solve_frontier_sharded) with world size 2
on the gloo backend.  The per-rank solver is a CPU stand-in with the frontier_* interface of capi.Solver -- a tiny
deterministic best-first tree search over integer "nodes" -- so what is checked here is the protocol: identical ramp-up,
fingerprint comparison, split by hash, min-all-reduce of the incumbent objectives, termination, winner / bound combination.
The device search behind the same calls is covered by tests/test_gpu_frontier.py."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ToyInfo:
    pass


def make_toy():
    sys.path.insert(0, ROOT)
    import planner_miqp_b200  # noqa: F401
    from planner_miqp_b200.capi import SolveInfo

    class ToySolver:
        """Plan k = binary tree of depth 10; leaf value = hash; node bound = min over its leaves minus a margin.  One node per
        plan and round.  The optimum is min over leaves, whatever the partition."""
        DEPTH = 10

        def leaf(self, k, path):
            return float(((path * 2654435761 + k * 40503) >> 7) % 1000) + 10.0

        def lb(self, k, depth, path):
            lo, hi = path << (self.DEPTH - depth), ((path + 1) << (self.DEPTH - depth))
            return min(self.leaf(k, q) for q in range(lo, hi)) - 0.25 * (self.DEPTH - depth)

        def upload(self, problems, gap_tol=None, time_limit=None, warm=None):
            self.n = len(problems)
            self.problems = problems

        def frontier_start(self):
            self.open = [[(self.lb(k, 0, 0), 0, 0)] for k in range(self.n)]
            self.ub = np.full(self.n, np.inf)
            self.own = np.full(self.n, np.inf)
            self.ownpath = [-1] * self.n
            self.pruned = np.full(self.n, np.inf)
            self.nodes = [0] * self.n

        def frontier_rounds(self, nrounds):
            r = 0
            while nrounds < 0 or r < nrounds:
                busy = 0
                for k in range(self.n):
                    keep = [e for e in self.open[k] if e[0] < self.ub[k]]
                    for e in self.open[k]:
                        if e[0] >= self.ub[k]:
                            self.pruned[k] = min(self.pruned[k], e[0])
                    self.open[k] = sorted(keep)
                    if not self.open[k]:
                        continue
                    busy += 1
                    b, d, path = self.open[k].pop(0)
                    self.nodes[k] += 1
                    if d == self.DEPTH:
                        v = self.leaf(k, path)
                        if v < self.own[k]:
                            self.own[k], self.ownpath[k] = v, path
                        self.ub[k] = min(self.ub[k], v)
                    else:
                        for c in (0, 1):
                            self.open[k].append((self.lb(k, d + 1, 2 * path + c), d + 1, 2 * path + c))
                r += 1
                if busy == 0:
                    break
            return sum(1 for k in range(self.n) if any(e[0] < self.ub[k] for e in self.open[k]))

        def frontier_fingerprint(self):
            return np.array([sum(hash((d, p)) % 1000003 for _, d, p in self.open[k]) for k in range(self.n)], dtype=np.int64)

        def frontier_split(self, rank, world):
            for k in range(self.n):
                self.open[k] = [e for e in self.open[k] if (e[2] * 7 + e[1]) % world == rank]

        def frontier_get_ub(self):
            return self.ub.copy()

        def frontier_tighten(self, ub):
            self.ub = np.minimum(self.ub, ub)

        def frontier_finish(self):
            return 0.0

        def fetch(self):
            xs, infos = [], []
            for k in range(self.n):
                have = np.isfinite(self.own[k])
                x = np.zeros(4)
                if have:
                    x[:] = [self.ownpath[k], self.own[k], k, 1.0]
                bb = min([e[0] for e in self.open[k]] + [self.pruned[k]])
                bb = min(bb, self.ub[k])
                infos.append(SolveInfo(0 if have else 1, False, self.own[k] if have else float("nan"), bb, float("nan"), 0.0,
                                       0.0 if have else float("nan"), self.nodes[k], 0, 0))
                xs.append(x)
            return xs, infos

    return ToySolver


class _Plan:
    scal = {"relative_mip_gap_tolerance": 1e-9}


def _worker(rank, world, port, q, perturb=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Toy = make_toy()
    from planner_miqp_b200.sharding import solve_frontier_sharded
    st = {}
    toy = Toy()
    if perturb and rank == 1:      # a rank whose ramp-up went differently: the fingerprints disagree
        fp0 = toy.frontier_fingerprint
        toy.frontier_fingerprint = lambda: fp0() + 1
    xs, infos = solve_frontier_sharded(toy, [_Plan() for _ in range(3)], ramp_rounds=4, exchange_every=3, stats=st)
    q.put((rank, [i.objective for i in infos], [i.best_bound for i in infos], [x.tolist() for x in xs],
           [i.nodes for i in infos], [bool(i.proven) for i in infos], st))
    dist.barrier()
    dist.destroy_process_group()


def test_frontier_protocol_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Toy = make_toy()
    t = Toy()
    want = [min(t.leaf(k, path) for path in range(1 << t.DEPTH)) for k in range(3)]
    for r in res:
        assert r[1] == want                                   # global optimum on every rank
        assert all(abs(b - o) <= 1e-9 * abs(o) for b, o in zip(r[2], r[1])) and all(r[5])
        assert [x[1] for x in r[3]] == want and all(x[3] == 1.0 for x in r[3])   # the winner's vector, exactly once
        assert r[6]["split"] and r[6]["exchanges"] >= 1 and r[6]["world"] == 2
    assert res[0][1:5] == res[1][1:5]
    # the shards did different work, and sharing incumbents kept the total near the single-rank node count
    single = Toy()
    single.upload([_Plan()] * 3)
    single.frontier_start()
    single.frontier_rounds(-1)
    assert res[0][6]["nodes_this_rank"] != res[1][6]["nodes_this_rank"] or res[0][6]["nodes_this_rank"] > 0
    assert sum(res[0][4]) <= 3 * sum(single.nodes)


def test_diverged_ramp_up_is_not_split():
    """different open lists after the ramp-up: no rank may drop nodes (every rank searches everything; still the optimum)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = make_toy()()
    want = [min(t.leaf(k, path) for path in range(1 << t.DEPTH)) for k in range(3)]
    for r in res:
        assert not r[6]["split"]
        assert r[1] == want and all(r[5])


def test_frontier_single_process_is_plain_solve():
    Toy = make_toy()
    from planner_miqp_b200.sharding import solve_frontier_sharded
    st = {}
    xs, infos = solve_frontier_sharded(Toy(), [_Plan()], stats=st)
    t = Toy()
    assert infos[0].objective == min(t.leaf(0, path) for path in range(1 << t.DEPTH))
    assert st["world"] == 1 and not st["split"] and st["exchanges"] == 0
