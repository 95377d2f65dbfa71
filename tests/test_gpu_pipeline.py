"""GPU: the execution plan of a batched call (solvers._execute) -- chunks of the batch pipelined through two
contexts of a GPU, the batch cut over several GPUs from one process -- gives exactly the results of one plain
call: slices are independent solves and every slice stops on its own test.  Also: two Python threads sharing the
default context (ctypes releases the GIL during the solve) do not disturb each other."""
import threading

import numpy as np
import pytest

import xinvert_b200 as xb
from xinvert_b200 import solvers
from tests import cases
from tests.test_apps_host import _c1_zeta

pytestmark = pytest.mark.gpu
DA = xb.DataArray


def _poisson_batch(T=12, ny=60, nx=128):
    zeta, co = _c1_zeta(ny, nx)
    rng = np.random.default_rng(3)
    z = np.stack([zeta * (1 + 0.3 * t) + 1e-6 * rng.standard_normal((ny, nx)) for t in range(T)])
    lam, phi = np.deg2rad(co['lon'])[None, :], np.deg2rad(co['lat'])[:, None]
    z[:, np.sin(5 * lam) * np.cos(3 * phi) > 0.6] = np.nan
    return DA(z, ['time', 'lat', 'lon'], dict(co, time=np.arange(T)))


@pytest.mark.parametrize("devices", [[0], [0, 0]])
def test_pipelined_poisson_front_end_equals_plain_call(gpu_ctx, monkeypatch, devices):
    monkeypatch.setattr(solvers, "PIPE_MIN_CELLS", 2 * 60 * 128)     # chunks of two slices
    F = _poisson_batch()
    ip = {'BCs': ['extend', 'periodic'], 'tolerance': 1e-7, 'mxLoop': 2000, 'printInfo': False}
    ip1 = dict(ip, ctx=gpu_ctx)                                       # explicit context: one plain call
    s1 = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ip1)
    ip2 = dict(ip, devices=devices)
    s2 = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ip2)
    assert np.array_equal(s1.values, s2.values, equal_nan=True)
    # loop counts and overflow flags identical; the last relative change to 1e-9: the cluster engine adds the partial
    # sums of |S| in an order that depends on the cluster shape, which is chosen per call (here: 12 slices vs 2 per chunk)
    assert np.array_equal(ip1['flags_all'][:, [0, 2]], ip2['flags_all'][:, [0, 2]])
    assert np.allclose(ip1['flags_all'][:, 1], ip2['flags_all'][:, 1], rtol=1e-9, atol=0)
    assert len(set(ip1['flags_all'][:, 2])) > 1                       # the slices stop at different sweeps
    assert "pipeline" not in ip1['stats'] and ip2['stats']['pipeline']['chunks'] >= 4
    assert ip2['stats']['pipeline']['workers'] == 2
    assert ip2['stats']['cell_updates'] == ip1['stats']['cell_updates']


def test_pipelined_dense_general_form_and_3d(gpu_ctx, monkeypatch):
    monkeypatch.setattr(solvers, "PIPE_MIN_CELLS", 1)
    c = cases.random_gen2d_rowcoef(40, 64, seed=2, batch=6)
    p = c["p"]
    args = ("fixed", "periodic", p["del1"], p["del1Sqr"], p["ratio"], p["ratioQtr"], p["ratioSqr"], p["optArg"], cases.UNDEF)
    S1, S2 = c["S0"].copy(), c["S0"].copy()
    f1, st1 = xb.solve_general_2D(S1, c["A"], None, c["C"], c["D"], c["E"], c["F"], c["G"], *args, mxLoop=300, tolerance=1e-6, ctx=gpu_ctx)
    f2, st2 = xb.solve_general_2D(S2, c["A"], None, c["C"], c["D"], c["E"], c["F"], c["G"], *args, mxLoop=300, tolerance=1e-6)
    assert np.array_equal(S1, S2) and np.array_equal(f1, f2) and st2["pipeline"]["chunks"] == 2 * solvers.PIPE_STREAMS
    c = cases.random_std3d(6, 20, 64, seed=3, batch=4)
    p = c["p"]
    args = ("fixed", "extend", "periodic", p["del1Sqr"], p["ratio2Sqr"], p["ratio1Sqr"], p["optArg"], cases.UNDEF)
    S1, S2 = c["S0"].copy(), c["S0"].copy()
    f1, _ = xb.solve_standard_3D(S1, c["A"], c["B"], c["C"], c["F"], *args, mxLoop=50, tolerance=1e-6, ctx=gpu_ctx)
    f2, st2 = xb.solve_standard_3D(S2, c["A"], c["B"], c["C"], c["F"], *args, mxLoop=50, tolerance=1e-6, devices=[0, 0])
    assert np.array_equal(S1, S2) and np.array_equal(f1, f2) and st2["pipeline"]["workers"] == 2


@pytest.mark.skipif(xb.device_count() < 2, reason="needs two GPUs")
def test_batch_cut_over_two_gpus_equals_one_gpu(gpu_ctx):
    """iParams['devices'] = [0, 1] from ONE process: standard 2-D (front end and dense), general 2-D, standard 3-D."""
    F = _poisson_batch(T=9)
    ip = {'BCs': ['fixed', 'periodic'], 'tolerance': 1e-7, 'mxLoop': 2000, 'printInfo': False}
    ip1, ip2 = dict(ip, ctx=gpu_ctx), dict(ip, devices=[0, 1])
    s1 = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ip1)
    s2 = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ip2)
    assert np.array_equal(s1.values, s2.values, equal_nan=True) and np.array_equal(ip1['flags_all'], ip2['flags_all'])
    assert ip2['stats']['pipeline']['devices'] == [0, 1]
    c = cases.random_std2d(40, 64, with_B=True, seed=1, batch=5)     # 9-point: colour engine
    p = c["p"]
    args = ("fixed", "fixed", p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], 1.2, cases.UNDEF)
    S1, S2 = c["S0"].copy(), c["S0"].copy()
    f1, _ = xb.solve_standard_2D(S1, c["A"], c["B"], c["C"], c["F"], *args, mxLoop=40, tolerance=-1.0, ctx=gpu_ctx)
    f2, _ = xb.solve_standard_2D(S2, c["A"], c["B"], c["C"], c["F"], *args, mxLoop=40, tolerance=-1.0, devices=[0, 1])
    assert np.array_equal(S1, S2) and np.array_equal(f1, f2)
    c = cases.random_gen2d_rowcoef(40, 64, seed=2, batch=6)
    p = c["p"]
    args = ("fixed", "periodic", p["del1"], p["del1Sqr"], p["ratio"], p["ratioQtr"], p["ratioSqr"], p["optArg"], cases.UNDEF)
    S1, S2 = c["S0"].copy(), c["S0"].copy()
    xb.solve_general_2D(S1, c["A"], None, c["C"], c["D"], c["E"], c["F"], c["G"], *args, mxLoop=100, tolerance=-1.0, ctx=gpu_ctx)
    xb.solve_general_2D(S2, c["A"], None, c["C"], c["D"], c["E"], c["F"], c["G"], *args, mxLoop=100, tolerance=-1.0, devices=[1, 0])
    assert np.array_equal(S1, S2)
    c = cases.random_std3d(6, 20, 64, seed=3, batch=4)
    p = c["p"]
    args = ("fixed", "extend", "periodic", p["del1Sqr"], p["ratio2Sqr"], p["ratio1Sqr"], p["optArg"], cases.UNDEF)
    S1, S2 = c["S0"].copy(), c["S0"].copy()
    xb.solve_standard_3D(S1, c["A"], c["B"], c["C"], c["F"], *args, mxLoop=30, tolerance=-1.0, ctx=gpu_ctx)
    xb.solve_standard_3D(S2, c["A"], c["B"], c["C"], c["F"], *args, mxLoop=30, tolerance=-1.0, devices=[0, 1])
    assert np.array_equal(S1, S2)


def test_two_python_threads_share_the_default_context(gpu_ctx):
    """The facade serialises begin..end sequences on a context (Context.lock): concurrent calls from two threads (the
    dask threaded scheduler does that) both get the right answer."""
    import oracle
    cs = [cases.poisson_latlon(60, 128, land=True, noise=1e-6, seed=s) for s in (1, 2)]
    want = [cases.run_std2d(oracle, c, "extend", "periodic", 400, 1e-7, omega=1.4, ordering="colour") for c in cs]
    got = [None, None]

    def work(k):
        for _ in range(5):
            got[k] = cases.run_std2d(xb, cs[k], "extend", "periodic", 400, 1e-7, omega=1.4)

    ts = [threading.Thread(target=work, args=(k,)) for k in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for k in (0, 1):
        assert np.array_equal(got[k][0], want[k][0]) and got[k][1][2] == want[k][1][2]


def test_batch_beyond_65535_slices_is_cut_into_several_calls(gpu_ctx):
    """The reference's serial loop takes any number of slices; one C-ABI call takes 65535."""
    nb, ny, nx = 66000, 5, 8
    rng = np.random.default_rng(0)
    A = np.ones((ny, nx)); Cc = np.ones((ny, nx))
    F = 1e-3 * rng.standard_normal((nb, ny, nx))
    S = np.zeros((nb, ny, nx))
    fl, st = xb.solve_standard_2D(S, A, None, Cc, F, "fixed", "fixed", 1.0, 0.25, 1.0, 1.2, cases.UNDEF, mxLoop=30, tolerance=1e-9,
                                  ctx=gpu_ctx)
    assert fl.shape == (nb, 3) and (fl[:, 2] > 0).all()
    import oracle
    for b in (0, 65534, 65535, 65999):
        So = np.zeros((ny, nx)); fo = np.array([0.0, 1.0, 0.0])
        oracle.invert_standard_2D(So, A, None, Cc, np.ascontiguousarray(F[b]), ny, nx, 0.0, 0.0, "fixed", "fixed", 1.0, 0.25, 1.0,
                                  1.2, cases.UNDEF, fo, 30, 1e-9, ordering="colour")
        assert np.array_equal(S[b], So) and fl[b, 2] == fo[2]
