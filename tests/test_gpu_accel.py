"""GPU: the optional Chebyshev schedule of the relaxation factor (include/xinv.h, xinv_opts.accel;
iParams['accel'] = 'chebyshev') on the cluster, resident and colour engines, against the
oracle's colour ordering driven by the same schedule.

Bar: fields BIT-EXACT, identical loop counts; flags[1] to 1e-6 relative.  The mode is NOT in the
reference: what is checked is that the accelerated CUDA path computes what an independent CPU
restatement of the same schedule computes, and that it converges to the reference's fixed point.  (Measured on C1 at omega_opt, tol 1e-10: 653 sweeps with the
schedule against 622 without -- under the reference's stop test, the relative change of mean|S|, the schedule does
NOT save sweeps; DESIGN.md section 8 records this.)"""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases

pytestmark = pytest.mark.gpu

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]


def _check_flags(f_gpu, f_ref):
    assert f_gpu[0] == f_ref[0]
    assert f_gpu[2] == f_ref[2]
    assert np.isclose(f_gpu[1], f_ref[1], rtol=1e-6, atol=1e-13)


@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("engine", ["colour", "resident", "cluster"])
def test_std2d_chebyshev_bit_exact(gpu_ctx, bcy, bcx, engine):
    c = cases.random_std2d_rowcoef(33, 48, seed=4)
    for sweeps in (0, 1, 2, 9):
        S_o, f_o = cases.run_std2d(oracle, c, bcy, bcx, sweeps, -1.0, omega=1.6, ordering="chebyshev")
        S_g, f_g = cases.run_std2d(xb, c, bcy, bcx, sweeps, -1.0, omega=1.6, engine=engine, accel="chebyshev")
        assert gpu_ctx.stats()["engine"] == engine
        assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()}"
        _check_flags(f_g, f_o)
    S_p, _ = cases.run_std2d(oracle, c, bcy, bcx, 9, -1.0, omega=1.6, ordering="colour")
    assert not np.array_equal(S_p, S_o)            # the schedule really changes the iterates


@pytest.mark.parametrize("engine", ["colour", "resident"])
@pytest.mark.parametrize("with_B", [False, True])
def test_ninepoint_and_general_chebyshev_bit_exact(gpu_ctx, engine, with_B):
    """4-colour scheme: colours 0, 1 with omega_h, colours 2, 3 with omega_h+1; general form; odd nx + periodic-x."""
    c = cases.random_std2d(30, 41, with_B=with_B, seed=6)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "periodic", 7, -1.0, omega=1.5, ordering="chebyshev")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "periodic", 7, -1.0, omega=1.5, engine=engine, accel="chebyshev")
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)
    c = cases.random_gen2d(28, 40, with_B=with_B, seed=7)
    S_o, f_o = cases.run_gen2d(oracle, c, "extend", "fixed", 7, -1.0, omega=1.5, ordering="chebyshev")
    S_g, f_g = cases.run_gen2d(xb, c, "extend", "fixed", 7, -1.0, omega=1.5, engine=engine, accel="chebyshev")
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)


def test_gen2d_cluster_and_std3d_colour_chebyshev(gpu_ctx):
    c = cases.random_gen2d_rowcoef(40, 64, seed=8)
    S_o, f_o = cases.run_gen2d(oracle, c, "fixed", "periodic", 11, -1.0, omega=1.7, ordering="chebyshev")
    S_g, f_g = cases.run_gen2d(xb, c, "fixed", "periodic", 11, -1.0, omega=1.7, accel="chebyshev")
    assert gpu_ctx.stats()["engine"] == "cluster"
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)
    c = cases.random_std3d(7, 20, 32, seed=9)
    S_o, f_o = cases.run_std3d(oracle, c, "extend", "periodic", 6, -1.0, omega=1.4, ordering="chebyshev")
    S_g, f_g = cases.run_std3d(xb, c, "extend", "periodic", 6, -1.0, omega=1.4, accel="chebyshev")
    assert gpu_ctx.stats()["engine"] == "colour"
    assert np.array_equal(S_g, S_o)
    _check_flags(f_g, f_o)


def test_chebyshev_to_tolerance_same_fixed_point(gpu_ctx):
    """C1 (360 x 180, fixed/periodic) at the grid's omega_opt, solved to tolerance with and without the schedule: the
    same field (both within the tolerance of the fixed point), sweep counts within 10 % of each other; several
    launches (check_every) change nothing."""
    c = cases.poisson_latlon(180, 360, land=False, noise=0.0, seed=0)
    om = c["p"]["optArg"]
    S_p, f_p = cases.run_std2d(xb, c, "fixed", "periodic", 9000, 1e-10, omega=om)
    S_a, f_a = cases.run_std2d(xb, c, "fixed", "periodic", 9000, 1e-10, omega=om, accel="chebyshev")
    assert gpu_ctx.stats()["engine"] in ("cluster", "colour")      # (colour where the device cannot host a 16-CTA cluster)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "periodic", 9000, 1e-10, omega=om, ordering="chebyshev")
    assert np.array_equal(S_a, S_o)
    _check_flags(f_a, f_o)
    assert abs(f_a[2] - f_p[2]) <= 0.1 * f_p[2], (f_a, f_p)
    assert np.abs(S_a - S_p).max() <= 1e-6 * np.abs(S_p).max()
    S_k, f_k = cases.run_std2d(xb, c, "fixed", "periodic", 9000, 1e-10, omega=om, accel="chebyshev", check_every=37)
    assert np.array_equal(S_k, S_o) and f_k[2] == f_o[2]


def test_accel_refused_where_it_does_not_apply(gpu_ctx):
    c = cases.random_std2d_rowcoef(33, 48, seed=4)
    with pytest.raises(xb.XinvError):
        cases.run_std2d(xb, c, "fixed", "fixed", 3, -1.0, omega=1.6, engine="fused", accel="chebyshev")
    with pytest.raises(xb.XinvError):
        cases.run_std2d(xb, c, "fixed", "fixed", 3, -1.0, omega=1.6, ordering="lexicographic", accel="chebyshev")
    with pytest.raises(xb.XinvError):
        cases.run_std2d(xb, c, "fixed", "fixed", 3, -1.0, omega=2.5, accel="chebyshev")
