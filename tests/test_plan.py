"""Host logic of the execution plan (solvers._plan / _execute): the flattened batch is cut over devices and, on
each device, into chunks pipelined through two contexts.  CPU only: fake contexts, a fake C call."""
import threading

import numpy as np
import pytest

from xinvert_b200 import solvers


class FakeCtx:
    def __init__(self, device, k):
        self.device, self.k, self.lock, self.calls = device, k, threading.RLock(), []


def _fake_contexts():
    made = {}

    def get(dev, n):
        return [made.setdefault((dev, k), FakeCtx(dev, k)) for k in range(n)]
    return get, made


def _stats(n):
    return dict(sweeps_launched=1, kernel_launches=2, cell_updates=n, solve_ms=1.0, h2d_ms=0.5, d2h_ms=0.5,
                h2d_bytes=8 * n, d2h_bytes=8 * n, dom_ms=0.0, dom_launches=0, engine="fused", ncolours=2,
                iters_per_pass=2, row_coeffs=1, sweep_ms=0.1, slow_strips=0)


@pytest.mark.parametrize("batch,cells,devices", [(32, 720 * 1440, None), (256, 720 * 1440, list(range(8))), (1, 6480000, None),
                                                 (70000, 73 * 144, [0]), (5, 1000, [0, 1, 2]), (3, 10 ** 7, [0, 1, 2, 3]),
                                                 (140000, 100, [0, 1])])
def test_plan_covers_every_slice_once(batch, cells, devices):
    items = solvers._plan(batch, cells, devices, pipelined=True)
    seen = np.zeros(batch, dtype=int)
    for dev, w, lo, hi in items:
        assert 0 <= lo < hi <= batch and hi - lo <= solvers.MAX_BATCH and 0 <= w < solvers.PIPE_STREAMS
        seen[lo:hi] += 1
    assert (seen == 1).all()
    # devices own contiguous, ordered blocks (distributed.shard_bounds)
    devs = devices or [0]
    from xinvert_b200.distributed import shard_bounds
    for r, d in enumerate(devs):
        lo, hi = shard_bounds(batch, len(devs), r)
        mine = sorted((a, b) for dd, _, a, b in items if dd == d and a >= lo and b <= hi)
        if hi > lo and len(set(devs)) == len(devs):
            assert mine[0][0] == lo and mine[-1][1] == hi


def test_plan_does_not_pipeline_device_resident_operands_or_small_jobs():
    assert solvers._plan(32, 720 * 1440, None, pipelined=False) == [(0, 0, 0, 32)]
    assert solvers._plan(4, 1000, None, pipelined=True) == [(0, 0, 0, 4)]


def test_execute_runs_every_chunk_on_its_context_and_merges_stats():
    get, made = _fake_contexts()
    done = []

    def call(c, lo, hi):
        c.calls.append((lo, hi))
        done.append((c.device, c.k, lo, hi))
        return _stats(hi - lo)

    st = solvers._execute(call, 256, 720 * 1440, None, [0, 1, 2, 3], host=True, _contexts=get)
    seen = np.zeros(256, dtype=int)
    for dev, k, lo, hi in done:
        seen[lo:hi] += 1
        assert dev == lo // 64                                       # 64 slices per device, in order
    assert (seen == 1).all()
    assert st["cell_updates"] == 256 and st["pipeline"]["devices"] == [0, 1, 2, 3] and st["pipeline"]["workers"] == 8
    assert {k for (_, k) in made} == {0, 1}                          # two contexts per device


def test_execute_with_explicit_context_stays_on_it():
    ctx = FakeCtx(0, 0)
    st = solvers._execute(lambda c, lo, hi: (c.calls.append((lo, hi)), _stats(hi - lo))[1], 70000, 100, ctx, None, host=True)
    assert ctx.calls == [(0, 65535), (65535, 70000)] and st["cell_updates"] == 70000 and "pipeline" not in st


def test_execute_propagates_errors_from_workers():
    get, _ = _fake_contexts()

    def call(c, lo, hi):
        if lo > 0:
            raise RuntimeError("boom")
        return _stats(hi - lo)

    with pytest.raises(RuntimeError, match="boom"):
        solvers._execute(call, 64, 720 * 1440, None, [0, 1], host=True, _contexts=get)
