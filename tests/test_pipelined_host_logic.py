"""Host logic of capi.PipelinedSolver (batches in flight) without a GPU: job k goes to instance k mod depth, every instance works
through its jobs in order on its own thread, results come back in job order, the first error is raised after all threads ended.
The solver instances are stand-ins; the device side is covered by tests/test_gpu_pipelined.py."""
import threading
import time

import pytest

import planner_miqp_b200 as P


class FakeSolver:
    def __init__(self, name, log, fail_on=None):
        self.name, self.log, self.fail_on = name, log, fail_on

    def solve_prepared(self, b):
        if b == self.fail_on:
            raise RuntimeError(f"boom on {b}")
        self.log.append((self.name, b, threading.get_ident()))
        time.sleep(0.01)
        return ("x", b), ("info", self.name)

    def solve_prepared_compact(self, b):
        return self.solve_prepared(b)

    def run(self):
        self.log.append((self.name, "run", threading.get_ident()))
        return 1.0

    def close(self):
        self.log.append((self.name, "closed", 0))


def make(depth, fail_on=None):
    log = []
    pipe = P.PipelinedSolver.__new__(P.PipelinedSolver)      # no device: the instances are stand-ins
    pipe.solvers = [FakeSolver(f"s{w}", log, fail_on) for w in range(depth)]
    return pipe, log


def test_jobs_alternate_over_the_instances_and_results_keep_job_order():
    pipe, log = make(3)
    out = pipe.solve_stream(list(range(8)), stagger_s=0.005)
    assert [o[0][1] for o in out] == list(range(8))
    assert [o[1][1] for o in out] == [f"s{k % 3}" for k in range(8)]
    per_solver = {}
    for name, job, tid in log:
        per_solver.setdefault(name, []).append((job, tid))
    assert [j for j, _ in per_solver["s0"]] == [0, 3, 6] and [j for j, _ in per_solver["s1"]] == [1, 4, 7] and [j for j, _ in per_solver["s2"]] == [2, 5]
    assert len({tid for jobs in per_solver.values() for _, tid in jobs}) == 3     # one host thread per instance
    assert all(len({tid for _, tid in jobs}) == 1 for jobs in per_solver.values())


def test_resident_runs_and_depth_one():
    pipe, log = make(2)
    assert pipe.run_resident(5) == [1.0] * 5
    assert [n for n, j, _ in log if j == "run"].count("s0") == 3
    single, log1 = make(1)
    assert [o[0][1] for o in single.solve_stream_compact(["a", "b"])] == ["a", "b"]
    single.close()
    assert log1[-1][1] == "closed"


def test_first_error_is_raised_after_the_threads_ended():
    pipe, log = make(2, fail_on=3)
    with pytest.raises(RuntimeError, match="boom on 3"):
        pipe.solve_stream(list(range(6)))
    # the other instance finished its own jobs
    assert [j for n, j, _ in log if n == "s0"] == [0, 2, 4]
