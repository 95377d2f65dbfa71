"""GPU: the CUDA path (through the C-ABI) against the committed golden vectors
of the REFERENCE itself (tests/golden/*.npz; generator tests/golden/make_golden.py).

* XINV_ORDER_LEX reproduces the reference's lexicographic trajectory: fields
  BIT-EXACT, loop counts identical; flags[1] to 1e-6 relative (the device sums
  |S| with a fixed tree, the reference serially).
* The colour ordering is compared with the "bridge" fixtures: red-black /
  4-colour iterations carried out by the reference's own code through
  alternating masked one-sweep calls: fields BIT-EXACT.
"""
import numpy as np
import pytest

import xinvert_b200 as xb
from tests import cases, golden_io

pytestmark = pytest.mark.gpu

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]


def _flags(f, g):
    assert f[0] == g[0] and f[2] == g[2], (f, g)
    assert np.isclose(f[1], g[1], rtol=1e-6, atol=1e-18), (f, g)


@pytest.mark.parametrize("with_B", [0, 1])
@pytest.mark.parametrize("bcy,bcx", BCS)
def test_std2d_lex_golden(gpu_ctx, with_B, bcy, bcx):
    if with_B and bcx == "periodic":
        pytest.skip("9-point lexicographic with periodic-x is refused on the GPU (serial dependency)")
    c, out = golden_io.load(f"std2d_B{with_B}")
    for sweeps in (0, 9):
        S, fl = cases.run_std2d(xb, c, bcy, bcx, sweeps, -1.0, omega=1.4 if not with_B else 1.2,
                                ordering="lexicographic")
        assert np.array_equal(S, out[f"S_{bcy}_{bcx}_{sweeps}"])
        _flags(fl, out[f"fl_{bcy}_{bcx}_{sweeps}"])


@pytest.mark.parametrize("with_B", [0, 1])
@pytest.mark.parametrize("bcy,bcx", BCS)
def test_gen2d_lex_golden(gpu_ctx, with_B, bcy, bcx):
    if with_B and bcx == "periodic":
        pytest.skip("9-point lexicographic with periodic-x is refused on the GPU")
    c, out = golden_io.load(f"gen2d_B{with_B}")
    S, fl = cases.run_gen2d(xb, c, bcy, bcx, 9, -1.0, omega=1.3, ordering="lexicographic")
    assert np.array_equal(S, out[f"S_{bcy}_{bcx}_9"])
    _flags(fl, out[f"fl_{bcy}_{bcx}_9"])


@pytest.mark.parametrize("bcy,bcx", BCS)
def test_std3d_lex_golden(gpu_ctx, bcy, bcx):
    c, out = golden_io.load("std3d")
    S, fl = cases.run_std3d(xb, c, bcy, bcx, 7, -1.0, omega=1.3, ordering="lexicographic")
    assert np.array_equal(S, out[f"S_{bcy}_{bcx}_7"])
    _flags(fl, out[f"fl_{bcy}_{bcx}_7"])


@pytest.mark.parametrize("bcy,bcx", [("fixed", "periodic"), ("extend", "periodic")])
def test_poisson_to_tolerance_lex_golden(gpu_ctx, bcy, bcx):
    """Same loop count and bit-equal field as the reference's own solve to tol 1e-8."""
    c, out = golden_io.load("poisson_tol")
    S, fl = cases.run_std2d(xb, c, bcy, bcx, 5000, 1e-8, omega=1.4, ordering="lexicographic")
    _flags(fl, out[f"fl_{bcy}_{bcx}"])
    assert np.array_equal(S, out[f"S_{bcy}_{bcx}"])


def test_overflow_lex_golden(gpu_ctx):
    c, out = golden_io.load("overflow")
    S, fl = cases.run_std2d(xb, c, "fixed", "fixed", 5000, 1e-12, omega=7.0, ordering="lexicographic")
    assert fl[0] == 1.0 and fl[2] == out["fl"][2]
    assert np.array_equal(S, out["S"])


@pytest.mark.parametrize("engine", ["colour", "fused"])
@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_redblack_std2d_bridge_golden(gpu_ctx, bcx, engine):
    c, out = golden_io.load("bridge_std2d")
    S, _ = cases.run_std2d(xb, c, "fixed", bcx, 5, -1.0, omega=1.4, engine=engine)
    assert np.array_equal(S, out[f"S_{bcx}"])


@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_redblack_gen2d_bridge_golden(gpu_ctx, bcx):
    c, out = golden_io.load("bridge_gen2d")
    S, _ = cases.run_gen2d(xb, c, "fixed", bcx, 5, -1.0, omega=1.3)
    assert np.array_equal(S, out[f"S_{bcx}"])


@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_redblack_std3d_bridge_golden(gpu_ctx, bcx):
    c, out = golden_io.load("bridge_std3d")
    S, _ = cases.run_std3d(xb, c, "fixed", bcx, 4, -1.0, omega=1.3)
    assert np.array_equal(S, out[f"S_{bcx}"])


@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_fourcolour_std2d_bridge_golden(gpu_ctx, bcx):
    c, out = golden_io.load("bridge_std2d_9pt")
    S, _ = cases.run_std2d(xb, c, "fixed", bcx, 4, -1.0, omega=1.2)
    assert np.array_equal(S, out[f"S_{bcx}"])


# ---- the bridge with BCy = 'extend' (row copy applied outside the reference calls), several strips / tiles ----
@pytest.mark.parametrize("engine", ["colour", "fused"])
@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_redblack_std2d_extend_bridge_golden(gpu_ctx, bcx, engine):
    c, out = golden_io.load("bridge_std2d_extend")
    S, _ = cases.run_std2d(xb, c, "extend", bcx, 5, -1.0, omega=1.4, engine=engine)
    assert np.array_equal(S, out[f"S_{bcx}"])


@pytest.mark.parametrize("engine", ["colour", "fused"])
@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_redblack_std3d_extend_bridge_golden(gpu_ctx, bcx, engine):
    c, out = golden_io.load("bridge_std3d_extend")
    S, _ = cases.run_std3d(xb, c, "extend", bcx, 4, -1.0, omega=1.3, engine=engine)
    assert np.array_equal(S, out[f"S_{bcx}"])


@pytest.mark.parametrize("engine", ["colour", "fused"])
@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_redblack_std3d_bridge_golden_both_engines(gpu_ctx, bcx, engine):
    c, out = golden_io.load("bridge_std3d")
    S, _ = cases.run_std3d(xb, c, "fixed", bcx, 4, -1.0, omega=1.3, engine=engine)
    assert np.array_equal(S, out[f"S_{bcx}"])
