"""GPU parity of the TMA-fed warp-marching engine (XINV_ENGINE_FUSED; T = 1 or 2
red+black iterations per pass) against the ordering-matched C oracle: BIT-EXACT
fields, identical loop counts (same bar as tests/test_gpu_parity.py).

Two kernel families: the general one (A, C vary in x and y; XINV_FUSED_VARIANT picks
the instantiation: T, rows per TMA chunk, ring depth, CTAs/SM -- xinv_march2d.cuh:
XM_VARIANTS) and the RC one, chosen automatically when A and C are constant along x
(XINV_FUSED_RC_VARIANT; XINV_FUSED_RC=0 forces the general kernels).  XINV_FUSED_RB
forces the owned rows per strip so that small grids are cut into many strips."""
import os

import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases

pytestmark = pytest.mark.gpu

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]
SHAPES = [(40, 64), (33, 47), (3, 4), (28, 60), (29, 61), (57, 122), (130, 258), (200, 366)]


def _check(c, bcy, bcx, sweeps, tol=-1.0, omega=1.4, rc=None):
    S_o, f_o = cases.run_std2d(oracle, c, bcy, bcx, sweeps, tol, omega=omega, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, bcy, bcx, sweeps, tol, omega=omega, engine="fused")
    st = xb.default_context().stats()
    assert st["engine"] == "fused"
    if rc is not None:
        assert st["row_coeffs"] == int(rc)
    assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()} at {np.argwhere(S_g != S_o)[:5]}"
    assert f_g[0] == f_o[0] and f_g[2] == f_o[2]
    assert np.isclose(f_g[1], f_o[1], rtol=1e-6, atol=1e-13)    # a difference of two norms: tree sum (GPU) vs serial sum (oracle)


VARIANTS = ["0", "1", "2", "3", "4"]          # general kernels (xinv_march2d.cuh: XM_VARIANTS)
RC_VARIANTS = ["0", "1", "2", "3", "4", "5", "6", "7", "8"]  # RC kernels (XM_RC_VARIANTS; 2, 3, 5, 6, 7, 8 keep the records in shared memory; 6: T = 4; 7 (default), 8: 12 / 6 warps per CTA)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("shape", SHAPES)
def test_fused_bit_exact(gpu_ctx, monkeypatch, variant, bcy, bcx, shape):
    """General kernels: A and C vary in x and y."""
    monkeypatch.setenv("XINV_FUSED_VARIANT", variant)
    if bcx == "periodic" and shape[1] % 2:
        pytest.skip("odd nx + periodic-x uses the wrap-fix colours (colour engine)")
    c = cases.random_std2d(*shape, with_B=False, seed=shape[0] * 1000 + shape[1])
    for sweeps in (0, 1, 2, 6, 7):          # mxLoop: 1, 2, 3, 7, 8 sweeps (odd counts end a T=2 solve on a 1-iteration pass)
        _check(c, bcy, bcx, sweeps, rc=False)


@pytest.mark.parametrize("variant", RC_VARIANTS)
@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("shape", SHAPES)
def test_fused_rowcoef_bit_exact(gpu_ctx, monkeypatch, variant, bcy, bcx, shape):
    """RC kernels: A and C constant along x (incl. whole rows of undef coefficients)."""
    monkeypatch.setenv("XINV_FUSED_RC_VARIANT", variant)
    if bcx == "periodic" and shape[1] % 2:
        pytest.skip("odd nx + periodic-x uses the wrap-fix colours (colour engine)")
    c = cases.random_std2d_rowcoef(*shape, seed=shape[0] * 1000 + shape[1] + 1, undef_rows=True)
    for sweeps in (0, 1, 2, 6, 7):
        _check(c, bcy, bcx, sweeps, rc=True)


def test_fused_rowcoef_equals_general_kernels(gpu_ctx, monkeypatch):
    """The same row-constant problem through the general kernels (XINV_FUSED_RC=0)."""
    c = cases.random_std2d_rowcoef(57, 122, seed=4)
    _check(c, "extend", "periodic", 7, rc=True)
    monkeypatch.setenv("XINV_FUSED_RC", "0")
    _check(c, "extend", "periodic", 7, rc=False)


@pytest.mark.parametrize("variant", ["0", "2", "3", "4", "6"])
@pytest.mark.parametrize("rb", ["1", "3", "8", "17"])
@pytest.mark.parametrize("bcy,bcx", BCS)
def test_fused_many_strips(gpu_ctx, monkeypatch, variant, rb, bcy, bcx):
    """Force tiny strips: every strip boundary (rows and columns) falls inside the grid."""
    monkeypatch.setenv("XINV_FUSED_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED_RC_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED_RB", rb)
    for shape in [(41, 130), (64, 256)]:
        c = cases.random_std2d(*shape, with_B=False, seed=shape[0] + 7 * shape[1])
        r = cases.random_std2d_rowcoef(*shape, seed=shape[0] + 7 * shape[1] + 1)
        for sweeps in (0, 4, 5):
            _check(c, bcy, bcx, sweeps, rc=False)
            _check(r, bcy, bcx, sweeps, rc=True)


@pytest.mark.parametrize("rcflag", ["0", "1"])
@pytest.mark.parametrize("variant", VARIANTS + ["6"])
def test_fused_poisson_to_tolerance(gpu_ctx, monkeypatch, variant, rcflag):
    monkeypatch.setenv("XINV_FUSED_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED_RC_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED_RC", rcflag)
    c = cases.poisson_latlon(90, 180, land=True, noise=1e-6, seed=0)
    _check(c, "extend", "periodic", 5000, tol=1e-8, rc=(rcflag == "1"))
    _check(c, "fixed", "periodic", 5000, tol=1e-8, rc=(rcflag == "1"))


@pytest.mark.parametrize("rcflag", ["0", "1"])
@pytest.mark.parametrize("variant", ["1", "2", "3", "6"])
def test_fused_t2_redo_when_stopping_mid_pass(gpu_ctx, monkeypatch, variant, rcflag):
    """T = 2: tolerances chosen so that the stop test fires after the 1st and after the
    2nd iteration of a pass (even and odd sweep counts); the overshoot is rolled back."""
    monkeypatch.setenv("XINV_FUSED_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED_RC_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED_RC", rcflag)
    c = cases.poisson_latlon(60, 120, land=True, noise=1e-6, seed=5)
    seen = set()
    for tol in (3e-3, 2e-3, 1e-3, 7e-4, 5e-4, 3e-4, 2e-4, 1e-4, 5e-5):
        S_o, f_o = cases.run_std2d(oracle, c, "fixed", "periodic", 5000, tol, omega=1.4, ordering="colour")
        S_g, f_g = cases.run_std2d(xb, c, "fixed", "periodic", 5000, tol, omega=1.4, engine="fused")
        assert f_g[2] == f_o[2] and np.array_equal(S_g, S_o), (tol, f_g, f_o)
        seen.add(int(f_o[2]) & 1)
    assert seen == {0, 1}
    # overflow in the first iteration of a pass
    c = cases.random_std2d(30, 40, with_B=False, seed=9)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 5000, 1e-12, omega=7.0, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 5000, 1e-12, omega=7.0, engine="fused")
    assert f_o[0] == 1 and f_g[0] == 1 and f_g[2] == f_o[2]
    assert np.array_equal(S_g, S_o, equal_nan=True)


def test_fused_odd_nx_periodic_falls_back_to_colour_engine(gpu_ctx):
    c = cases.random_std2d(20, 31, with_B=False, seed=1)
    with pytest.raises(xb.XinvError):
        cases.run_std2d(xb, c, "fixed", "periodic", 3, -1.0, engine="fused")
    S_o, _ = cases.run_std2d(oracle, c, "fixed", "periodic", 3, -1.0, ordering="colour")
    S_g, _ = cases.run_std2d(xb, c, "fixed", "periodic", 3, -1.0, engine="auto")
    assert np.array_equal(S_g, S_o)
    assert xb.default_context().stats()["engine"] == "resident"      # small enough for one SM's shared memory
    c = cases.random_std2d(120, 131, with_B=False, seed=1)
    S_o, _ = cases.run_std2d(oracle, c, "fixed", "periodic", 3, -1.0, ordering="colour")
    S_g, _ = cases.run_std2d(xb, c, "fixed", "periodic", 3, -1.0, engine="auto")
    assert np.array_equal(S_g, S_o)
    assert xb.default_context().stats()["engine"] == "colour"


@pytest.mark.parametrize("rcflag", ["0", "1"])
@pytest.mark.parametrize("shared", [True, False])
def test_fused_batched_freeze(gpu_ctx, monkeypatch, shared, rcflag):
    """Batch of slices with shared (stride 0) or per-slice coefficients; every slice
    stops on its own test and equals its single-slice oracle run."""
    monkeypatch.setenv("XINV_FUSED_RC", rcflag)
    B = 4
    c = cases.poisson_latlon(60, 124, land=True, noise=1e-6, seed=2, batch=B)
    for b in range(B):
        c["F"][b][c["F"][b] != cases.UNDEF] *= (1.0 + 2.0 * b)
    p = c["p"]
    A = c["A"] if shared else np.ascontiguousarray(np.broadcast_to(c["A"], (B,) + c["A"].shape))
    Cc = c["C"] if shared else np.ascontiguousarray(np.broadcast_to(c["C"], (B,) + c["C"].shape))
    S = c["S0"].copy()
    fl, st = xb.solve_standard_2D(S, A, None, Cc, c["F"], "extend", "periodic", p["del1Sqr"], p["ratioQtr"],
                                  p["ratioSqr"], 1.4, mxLoop=3000, tolerance=1e-7, engine="fused")
    assert st["engine"] == "fused"
    for b in range(B):
        cb = dict(A=c["A"], C=c["C"], F=c["F"][b], S0=c["S0"][b], p=p)
        S_o, f_o = cases.run_std2d(oracle, cb, "extend", "periodic", 3000, 1e-7, omega=1.4, ordering="colour")
        assert fl[b, 2] == f_o[2]
        assert np.array_equal(S[b], S_o)
    assert len({int(x) for x in fl[:, 2]}) > 1          # they really stopped at different sweeps


def test_fused_device_pointers(gpu_ctx):
    """mem_space = DEVICE: operands are CUDA tensors, S updated in place."""
    import torch
    c = cases.poisson_latlon(64, 120, land=True, noise=1e-6, seed=3)
    p = c["p"]
    dev = torch.device("cuda", 0)
    S = torch.zeros(c["S0"].shape, dtype=torch.float64, device=dev)
    A, Cc, F = (torch.from_numpy(c[k]).to(dev) for k in ("A", "C", "F"))
    torch.cuda.synchronize()
    fl, st = xb.solve_standard_2D(S, A, None, Cc, F, "fixed", "periodic", p["del1Sqr"], p["ratioQtr"],
                                  p["ratioSqr"], 1.4, mxLoop=50, tolerance=-1.0, engine="fused")
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "periodic", 50, -1.0, omega=1.4, ordering="colour")
    assert st["h2d_bytes"] == 0 and st["d2h_bytes"] == 0
    assert np.array_equal(S.cpu().numpy(), S_o)


@pytest.mark.parametrize("variant", ["1", "2"])
def test_fused_extreme_denominators(gpu_ctx, monkeypatch, variant):
    """Coefficients of order 1e306 put the denominators of optArg / den at the edge of the
    double range (quotients near the denormal boundary): the precomputed factor array must
    still be the correctly rounded quotient the oracle forms in every sweep."""
    monkeypatch.setenv("XINV_FUSED_VARIANT", variant)
    c = cases.random_std2d(40, 64, with_B=False, seed=77)
    _check(c, "fixed", "periodic", 6)
    big = dict(c, A=c["A"] * 1e306, C=c["C"] * 1e306, F=np.where(c["F"] == cases.UNDEF, cases.UNDEF, c["F"] * 1e306))
    S_o, f_o = cases.run_std2d(oracle, big, "fixed", "periodic", 6, -1.0, omega=1.4, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, big, "fixed", "periodic", 6, -1.0, omega=1.4, engine="fused")
    assert np.isfinite(S_o).all()
    assert np.array_equal(S_g, S_o) and f_g[2] == f_o[2]
