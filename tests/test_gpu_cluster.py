"""GPU: XINV_ENGINE_CLUSTER (xinv_cluster2d.cuh) -- the whole solve of a 2-D slice with row
coefficients inside one thread-block cluster, psi in registers + distributed shared memory --
against the ordering-matched C oracle.

Bar: fields BIT-EXACT (np.array_equal), identical loop counts and overflow flags; flags[1] to 1e-6
relative (the cluster sums |S| thread -> warp -> cluster in a fixed order, the oracle serially).
Every cluster size R (1, 2, 4, 8, 16 CTAs) and every run length K (4, 6, 8, 10, 12, 16 cells per thread)
is forced through XINV_CLUSTER_R / XINV_CLUSTER_K.
"""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases
from xinvert_b200 import solvers

pytestmark = pytest.mark.gpu

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]


def _check_flags(f_gpu, f_ref):
    assert f_gpu[0] == f_ref[0]
    assert f_gpu[2] == f_ref[2]
    assert np.isclose(f_gpu[1], f_ref[1], rtol=1e-6, atol=1e-13)


def _engine(ctx):
    return ctx.stats()["engine"]


_R16 = None


def _r16_available():
    """Clusters of 16 CTAs are a non-portable size: whether one can be resident depends on the GPCs of the device."""
    global _R16
    if _R16 is None:
        import os
        old = os.environ.get("XINV_CLUSTER_R")
        os.environ["XINV_CLUSTER_R"] = "16"
        try:
            cases.run_std2d(xb, cases.random_std2d_rowcoef(50, 48, seed=1), "fixed", "fixed", 1, -1.0, engine="cluster")
            _R16 = True
        except xb.XinvError:
            _R16 = False
        finally:
            if old is None:
                os.environ.pop("XINV_CLUSTER_R", None)
            else:
                os.environ["XINV_CLUSTER_R"] = old
    return _R16


def _force(monkeypatch, R, K):
    if R == 16 and not _r16_available():
        pytest.skip("this device cannot host a cluster of 16 CTAs")
    monkeypatch.setenv("XINV_CLUSTER_R", str(R))
    monkeypatch.setenv("XINV_CLUSTER_K", str(K))


@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("R", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("K", [4, 6, 8, 10, 12, 16])
def test_std2d_cluster_bit_exact(gpu_ctx, monkeypatch, bcy, bcx, R, K):
    """Standard form, every (R, K): ragged nx (a partly filled last run) with fixed-x, nx a multiple of K
    with periodic-x; land cells, whole rows of undef coefficients."""
    _force(monkeypatch, R, K)
    ny = {1: 9, 2: 12, 4: 19, 8: 33, 16: 50}[R]
    nx = 4 * K * 3 if bcx == "periodic" else 4 * K * 3 + 5
    c = cases.random_std2d_rowcoef(ny, nx, seed=R * 100 + K, undef_rows=(R >= 8))
    for sweeps in (0, 1, 6):
        S_o, f_o = cases.run_std2d(oracle, c, bcy, bcx, sweeps, -1.0, omega=1.4, ordering="colour")
        S_g, f_g = cases.run_std2d(xb, c, bcy, bcx, sweeps, -1.0, omega=1.4, engine="cluster")
        assert _engine(gpu_ctx) == "cluster"
        assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()}"
        _check_flags(f_g, f_o)


@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("R,K", [(1, 4), (2, 8), (4, 12), (8, 6), (16, 16), (16, 12)])
def test_gen2d_cluster_bit_exact(gpu_ctx, monkeypatch, bcy, bcx, R, K):
    _force(monkeypatch, R, K)
    ny = 3 * R + 4
    nx = 2 * K * 5 if bcx == "periodic" else 2 * K * 5 + 3
    c = cases.random_gen2d_rowcoef(ny, nx, seed=R + K, undef_rows=(R == 4))
    S_o, f_o = cases.run_gen2d(oracle, c, bcy, bcx, 9, -1.0, omega=1.4, ordering="colour")
    S_g, f_g = cases.run_gen2d(xb, c, bcy, bcx, 9, -1.0, omega=1.4, engine="cluster")
    assert _engine(gpu_ctx) == "cluster"
    assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()}"
    _check_flags(f_g, f_o)


@pytest.mark.parametrize("wide", [False, True])
def test_cluster_several_warps_per_row(gpu_ctx, monkeypatch, wide):
    """More than 32 runs per row: a row is spread over several warps (K = 4, nx = 360 -> 90 runs, 3 warps)."""
    _force(monkeypatch, 16 if wide else 8, 4)
    c = cases.random_std2d_rowcoef(40 if wide else 20, 360, seed=9)
    S_o, f_o = cases.run_std2d(oracle, c, "extend", "periodic", 5, -1.0, omega=1.4, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "extend", "periodic", 5, -1.0, omega=1.4, engine="cluster")
    assert _engine(gpu_ctx) == "cluster"
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)


def test_cluster_c1_size_to_tolerance_auto(gpu_ctx):
    """BASELINE configs[0] (360 x 180 lat-lon Poisson, fixed/periodic, omega 1.4, tol 1e-8): engine='auto'
    picks the cluster engine; same loop count and bits as the oracle's colour ordering and as the
    marching engine; also through several launches (check_every)."""
    c = cases.poisson_latlon(180, 360, land=False, noise=0.0, seed=0)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "periodic", 5000, 1e-8, omega=1.4, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "periodic", 5000, 1e-8, omega=1.4)
    assert _engine(gpu_ctx) == ("cluster" if _r16_available() else "fused")      # 360 x 180 needs all 16 CTAs
    assert f_o[2] > 1000
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)
    S_m, f_m = cases.run_std2d(xb, c, "fixed", "periodic", 5000, 1e-8, omega=1.4, engine="fused")
    assert _engine(gpu_ctx) == "fused"
    assert np.array_equal(S_m, S_o) and f_m[2] == f_o[2]
    S_k, f_k = cases.run_std2d(xb, c, "fixed", "periodic", 5000, 1e-8, omega=1.4, check_every=100)
    assert np.array_equal(S_k, S_o) and f_k[2] == f_o[2]


def test_cluster_periodic_250_columns_auto(gpu_ctx):
    """A periodic grid whose width only K = 10 divides (250 columns): engine='auto' still finds a cluster shape."""
    c = cases.poisson_latlon(100, 250, land=True, noise=1e-6, seed=3)
    S_o, f_o = cases.run_std2d(oracle, c, "extend", "periodic", 40, -1.0, omega=1.5, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "extend", "periodic", 40, -1.0, omega=1.5)
    assert _engine(gpu_ctx) == "cluster"
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)


def test_cluster_land_mask_extend(gpu_ctx):
    """Helmholtz_ocean-style problem (tests/test_Poisson.py:49-53 of the reference): land mask, extend/periodic."""
    c = cases.poisson_latlon(90, 180, land=True, noise=1e-6, seed=0)
    S_o, f_o = cases.run_std2d(oracle, c, "extend", "periodic", 3000, 1e-9, omega=1.5, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "extend", "periodic", 3000, 1e-9, omega=1.5)
    assert _engine(gpu_ctx) == "cluster"
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)


@pytest.mark.parametrize("shared", [True, False])
def test_cluster_batch_per_slice_stop(gpu_ctx, shared):
    """More slices than clusters; shared (stride 0) or per-slice coefficients; every slice stops on its own test."""
    B = 40
    c = cases.poisson_latlon(46, 72, land=True, noise=1e-6, seed=2, batch=B)
    for b in range(B):
        c["F"][b][c["F"][b] != cases.UNDEF] *= (1.0 + 0.5 * b) ** 3
    p = c["p"]
    A = c["A"] if shared else np.ascontiguousarray(np.broadcast_to(c["A"], (B,) + c["A"].shape[-2:]))
    Cc = c["C"] if shared else np.ascontiguousarray(np.broadcast_to(c["C"], (B,) + c["C"].shape[-2:]))
    S = c["S0"].copy()
    fl, st = solvers.solve_standard_2D(S, A, None, Cc, c["F"], "extend", "periodic", p["del1Sqr"], p["ratioQtr"], p["ratioSqr"],
                                       1.3, cases.UNDEF, (0.0, 1.0, 0.0), 800, 1e-7)
    assert st["engine"] == "cluster"
    loops = set()
    for b in range(0, B, 3):
        cb = dict(c, F=c["F"][b], S0=c["S0"][b], A=c["A"] if c["A"].ndim == 2 else c["A"][b],
                  C=c["C"] if c["C"].ndim == 2 else c["C"][b])
        S_o, f_o = cases.run_std2d(oracle, cb, "extend", "periodic", 800, 1e-7, omega=1.3, ordering="colour")
        assert np.array_equal(S[b], S_o)
        _check_flags(fl[b], f_o)
        loops.add(int(f_o[2]))
    assert len(loops) > 3


def test_cluster_overflow_and_zero_exit(gpu_ctx):
    c = cases.random_std2d_rowcoef(20, 32, seed=3, land=0.0)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 5000, 1e-12, omega=7.0, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 5000, 1e-12, omega=7.0, engine="cluster")
    assert f_o[0] == 1.0 and f_g[0] == 1.0 and f_g[2] == f_o[2]
    assert np.array_equal(S_g, S_o, equal_nan=True)
    c["F"][:] = 0.0
    c["S0"][:] = 0.0                          # norm == 0 exit (numbas.py:410), 2-D standard form only
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 50, 1e-12, omega=1.4, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 50, 1e-12, omega=1.4, engine="cluster")
    assert f_g[2] == f_o[2] == 0 and np.array_equal(S_g, S_o)


def test_cluster_refused_where_it_does_not_apply(gpu_ctx):
    """x-varying coefficients / a slice too large for a cluster: engine='cluster' raises, 'auto' uses the marching engine."""
    c = cases.random_std2d(40, 64, with_B=False, seed=1)
    with pytest.raises(xb.XinvError):
        cases.run_std2d(xb, c, "fixed", "fixed", 3, -1.0, engine="cluster")
    big = cases.random_std2d_rowcoef(700, 1440, seed=2)
    with pytest.raises(xb.XinvError):
        cases.run_std2d(xb, big, "fixed", "fixed", 3, -1.0, engine="cluster")
    S_o, _ = cases.run_std2d(oracle, big, "fixed", "fixed", 3, -1.0, ordering="colour")
    S_g, _ = cases.run_std2d(xb, big, "fixed", "fixed", 3, -1.0)
    assert _engine(gpu_ctx) == "fused" and np.array_equal(S_g, S_o)
