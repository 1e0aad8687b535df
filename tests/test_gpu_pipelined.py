"""capi.PipelinedSolver: several solver instances of one GPU fed alternately by host threads (batches in flight).  Every batch is
still one deterministic search, so the results must equal those of a single Solver, whatever overlaps on the device."""
import numpy as np
import pytest

import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario

pytestmark = pytest.mark.gpu


def _batches(n_batches=5, size=48):
    return [[obstacle_scenario(100 * b + k).build() for k in range(size)] for b in range(n_batches)]


def test_batches_in_flight_equal_one_batch_at_a_time():
    batches = _batches()
    ref = P.Solver()
    want = [ref.solve_batch(ps, gap_tol=1e-4, time_limit=60) for ps in batches]
    ref.close()
    pipe = P.PipelinedSolver(depth=3)
    prepared = [pipe.prepare(ps, gap_tol=1e-4, time_limit=60) for ps in batches]
    got = pipe.solve_stream(prepared, stagger_s=0.002)
    assert len(got) == len(batches)
    for (xs0, infos0), (xs1, infos1) in zip(want, got):
        for a, b, xa, xb in zip(infos0, infos1, xs0, xs1):
            assert a.status == b.status == 0 and a.proven and b.proven
            assert a.objective == b.objective and a.nodes == b.nodes
            assert np.array_equal(xa, xb)
    # compact results of the same stream: the trajectory blocks of the vectors above
    from planner_miqp_b200.results import block_views
    got_c = pipe.solve_stream_compact(prepared)
    for ps, (xs0, infos0), (tr, infos2) in zip(batches, want, got_c):
        for p, x, t, a, c in zip(ps, xs0, tr, infos0, infos2):
            assert c.objective == a.objective
            v = block_views(p, x)
            assert np.array_equal(t[:, :, 0], v["pos_x"]) and np.array_equal(t[:, :, 3], v["pos_y"])
    # resident batches, timed on the device over all streams
    pipe.upload_resident(prepared[:3])
    total_ms, per_run = pipe.timed_resident(6)
    assert len(per_run) == 6 and all(ms > 0 for ms in per_run)
    assert 0 < total_ms <= sum(per_run) + 1.0          # the runs overlap: the span is at most the sum of the run times
    assert total_ms >= max(per_run) - 1e-3
    pipe.close()
