"""Host logic of the multi-GPU path on CPU: world_size-2 gloo.

No GPU compute here: the C-ABI stepper is replaced by a mock whose slices
"converge" after a prescribed number of sweeps.  What is checked is what the
ranks must agree on: the block partition of the batch axis, that every rank
issues the same number of collectives (a rank whose slices are all frozen keeps
stepping until the *global* active count is 0), and that a shard solved by the
C oracle slice by slice equals the unsharded solve (slices are independent: no
halo, no data-path collective).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_bounds_partition_every_slice_once():
    from xinvert_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 31, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(n, world, r)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [np.subtract(*shard_bounds(n, world, r)[::-1]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


class _MockStepper:
    """Slices that stop after `need[b]` sweeps; step(k) runs up to k more sweeps."""

    def __init__(self, need):
        self.need = list(need)
        self.done = 0
        self.calls = 0

    def step(self, k):
        k = k or 4
        self.calls += 1
        self.done += k
        return sum(1 for n in self.need if n > self.done)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from xinvert_b200 import distributed as xd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ar = xd.TorchAllReduce()
        # rank 0's slices finish early, rank 1's late: rank 0 must keep issuing collectives
        need = [3, 5] if rank == 0 else [40, 17, 9]
        st = _MockStepper(need)
        chunks = xd.run_until_all_done(st.step, ar, sweeps_per_chunk=4)
        total = ar(len(need))

        # sharded oracle solve == the same slices of an unsharded solve
        import oracle
        from tests import cases
        c = cases.poisson_latlon(24, 48, land=True, noise=1e-6, seed=3, batch=5)
        lo, hi = xd.shard_bounds(5, world, rank)
        out = []
        for b in range(lo, hi):
            cc = dict(A=c["A"], C=c["C"], F=c["F"][b], S0=c["S0"][b], p=c["p"])
            S, fl = cases.run_std2d(oracle, cc, "fixed", "periodic", 50, 1e-6, omega=1.4, ordering="colour")
            out.append((b, S, fl))
        q.put((rank, chunks, st.calls, total, lo, hi, out))
    finally:
        dist.destroy_process_group()


def test_two_ranks_stay_in_lock_step_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r = q.get(timeout=180)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # 40 sweeps at 4 per chunk -> 10 chunks on BOTH ranks
    assert res[0][1] == res[1][1] == 10
    assert res[0][2] == res[1][2] == 10
    assert res[0][3] == res[1][3] == 5
    assert (res[0][4], res[0][5], res[1][4], res[1][5]) == (0, 3, 3, 5)

    import oracle
    from tests import cases
    c = cases.poisson_latlon(24, 48, land=True, noise=1e-6, seed=3, batch=5)
    got = {b: (S, fl) for r in (0, 1) for b, S, fl in res[r][6]}
    assert sorted(got) == [0, 1, 2, 3, 4]
    for b in range(5):
        cc = dict(A=c["A"], C=c["C"], F=c["F"][b], S0=c["S0"][b], p=c["p"])
        S, fl = cases.run_std2d(oracle, cc, "fixed", "periodic", 50, 1e-6, omega=1.4, ordering="colour")
        assert np.array_equal(S, got[b][0]) and np.array_equal(fl, got[b][1])


def test_run_until_all_done_single_rank_identity():
    from xinvert_b200 import distributed as xd
    st = _MockStepper([9, 2])
    assert xd.run_until_all_done(st.step, lambda v: v, sweeps_per_chunk=4) == 3
