"""GPU parity of the fused engine on the general form (invert_general_2D, numbas.py:987-1201)
with coefficients constant along x (Gill-Matsuno, Stommel) against the ordering-matched C
oracle: BIT-EXACT fields, identical loop counts.  Coefficients that vary along x stay on the
colour engine."""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases

pytestmark = pytest.mark.gpu

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]
SHAPES = [(40, 64), (33, 47), (3, 4), (29, 61), (57, 122), (130, 258), (200, 366)]


def _check(c, bcy, bcx, sweeps, tol=-1.0, omega=1.4, engine="fused"):
    S_o, f_o = cases.run_gen2d(oracle, c, bcy, bcx, sweeps, tol, omega=omega, ordering="colour")
    S_g, f_g = cases.run_gen2d(xb, c, bcy, bcx, sweeps, tol, omega=omega, engine=engine)
    st = xb.default_context().stats()
    assert st["engine"] in (("fused", "cluster") if engine == "auto" else ("fused",)) and st["row_coeffs"] == 1
    assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()} at {np.argwhere(S_g != S_o)[:5]}"
    assert f_g[0] == f_o[0] and f_g[2] == f_o[2]
    assert np.isclose(f_g[1], f_o[1], rtol=1e-6, atol=1e-13)    # a difference of two norms: tree sum (GPU) vs serial sum (oracle)


@pytest.mark.parametrize("variant", ["0", "1"])
@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("shape", SHAPES)
def test_fused_gen2d_bit_exact(gpu_ctx, monkeypatch, variant, bcy, bcx, shape):
    monkeypatch.setenv("XINV_FUSED_GEN_VARIANT", variant)
    if bcx == "periodic" and shape[1] % 2:
        pytest.skip("odd nx + periodic-x uses the wrap-fix colours (colour engine)")
    c = cases.random_gen2d_rowcoef(*shape, seed=shape[0] * 1000 + shape[1], undef_rows=True)
    for sweeps in (0, 1, 2, 6, 7):
        _check(c, bcy, bcx, sweeps)


@pytest.mark.parametrize("variant", ["0", "1"])
@pytest.mark.parametrize("rb", ["1", "3", "8", "17"])
def test_fused_gen2d_many_strips(gpu_ctx, monkeypatch, variant, rb):
    monkeypatch.setenv("XINV_FUSED_GEN_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED_RB", rb)
    for shape in [(41, 130), (64, 256)]:
        c = cases.random_gen2d_rowcoef(*shape, seed=shape[0] + 7 * shape[1])
        for bcy, bcx in BCS:
            _check(c, bcy, bcx, 5)


def test_fused_gen2d_to_tolerance_and_auto_engine(gpu_ctx):
    c = cases.random_gen2d_rowcoef(60, 120, seed=11)
    _check(c, "fixed", "periodic", 3000, tol=1e-9, engine="auto")
    _check(c, "extend", "fixed", 3000, tol=1e-9, engine="auto")


def test_gen2d_x_varying_coefficients_stay_on_colour_engine(gpu_ctx):
    c = cases.random_gen2d(80, 128, with_B=False, seed=5)
    with pytest.raises(xb.XinvError):
        cases.run_gen2d(xb, c, "fixed", "periodic", 3, -1.0, engine="fused")
    S_o, f_o = cases.run_gen2d(oracle, c, "fixed", "periodic", 3, -1.0, ordering="colour")
    S_g, f_g = cases.run_gen2d(xb, c, "fixed", "periodic", 3, -1.0, engine="auto")
    assert xb.default_context().stats()["engine"] == "colour"
    assert np.array_equal(S_g, S_o)


def test_fused_gen2d_batched_freeze(gpu_ctx):
    """Several slices of G over shared coefficients; each stops on its own test."""
    B = 3
    c = cases.random_gen2d_rowcoef(48, 96, seed=21, batch=B)
    one = cases.random_gen2d_rowcoef(48, 96, seed=21)            # same coefficient rows (same seed + offset)
    for k in "ACDEF":
        c[k] = one[k]
    for b in range(B):
        c["G"][b][c["G"][b] != cases.UNDEF] *= (1.0 + 3.0 * b)
    p = c["p"]
    S = c["S0"].copy()
    fl, st = xb.solve_general_2D(S, c["A"], None, c["C"], c["D"], c["E"], c["F"], c["G"], "fixed", "periodic",
                                 p["del1"], p["del1Sqr"], p["ratio"], p["ratioQtr"], p["ratioSqr"], 1.4,
                                 mxLoop=2000, tolerance=1e-7, engine="fused")
    assert st["engine"] == "fused"
    for b in range(B):
        cb = dict(c, G=c["G"][b], S0=c["S0"][b])
        S_o, f_o = cases.run_gen2d(oracle, cb, "fixed", "periodic", 2000, 1e-7, omega=1.4, ordering="colour")
        assert fl[b, 2] == f_o[2]
        assert np.array_equal(S[b], S_o)
