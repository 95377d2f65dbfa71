"""GPU: XINV_ENGINE_RESIDENT (xinv_resident.cuh) -- the whole solve of a small 2-D slice inside one
CTA, operands in shared memory -- against the ordering-matched C oracle and the reference's own
bridge goldens.

Bar: fields BIT-EXACT (np.array_equal), identical loop counts and overflow flags; flags[1] to 1e-6
relative (a block-tree sum of |S| against the oracle's serial one).
"""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases, golden_io
from xinvert_b200 import solvers

pytestmark = pytest.mark.gpu

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]


def _check_flags(f_gpu, f_ref):
    assert f_gpu[0] == f_ref[0]
    assert f_gpu[2] == f_ref[2]
    assert np.isclose(f_gpu[1], f_ref[1], rtol=1e-6, atol=1e-13)


def _engine(ctx):
    return ctx.stats()["engine"]


@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("shape", [(37, 73), (33, 47), (3, 3), (40, 64), (5, 4)])
@pytest.mark.parametrize("with_B", [False, True])
def test_std2d_resident_bit_exact(gpu_ctx, bcy, bcx, shape, with_B):
    """5-point (two colours) and 9-point (four colours; wrap-fix colours for odd nx + periodic-x)."""
    c = cases.random_std2d(*shape, with_B=with_B, seed=(shape[0] * 131 + shape[1] + with_B) % 1000)
    for sweeps in (0, 1, 7):
        S_o, f_o = cases.run_std2d(oracle, c, bcy, bcx, sweeps, -1.0, omega=1.4, ordering="colour")
        S_g, f_g = cases.run_std2d(xb, c, bcy, bcx, sweeps, -1.0, omega=1.4, engine="resident")
        assert _engine(gpu_ctx) == "resident"
        assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()}"
        _check_flags(f_g, f_o)


@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("with_B", [False, True])
def test_gen2d_resident_bit_exact(gpu_ctx, bcy, bcx, with_B):
    for shape in [(33, 47), (36, 50)]:
        c = cases.random_gen2d(*shape, with_B=with_B, seed=11)
        S_o, f_o = cases.run_gen2d(oracle, c, bcy, bcx, 9, -1.0, omega=1.4, ordering="colour")
        S_g, f_g = cases.run_gen2d(xb, c, bcy, bcx, 9, -1.0, omega=1.4, engine="resident")
        assert _engine(gpu_ctx) == "resident"
        assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()}"
        _check_flags(f_g, f_o)


@pytest.mark.parametrize("with_B", [False, True])
def test_resident_to_tolerance_many_chunks(gpu_ctx, with_B):
    """Solved to tolerance over several launches (check_every = 16 sweeps per launch: psi goes back
    to HBM and is staged again between launches): same loop count and bits as the oracle, and as
    one launch for the whole solve."""
    c = cases.random_std2d(37, 73, with_B=with_B, seed=77)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 3000, 1e-9, omega=1.3, ordering="colour")
    assert f_o[2] > 40                     # really several launches
    S_a, f_a = cases.run_std2d(xb, c, "fixed", "fixed", 3000, 1e-9, omega=1.3, engine="resident", check_every=16)
    S_b, f_b = cases.run_std2d(xb, c, "fixed", "fixed", 3000, 1e-9, omega=1.3, engine="resident")
    for S_g, f_g in ((S_a, f_a), (S_b, f_b)):
        assert np.array_equal(S_g, S_o)
        _check_flags(f_g, f_o)


def test_resident_batch_per_slice_stop(gpu_ctx):
    """A batch: one slice per CTA, more slices than CTAs, every slice stops on its own test."""
    batch = 333
    c = cases.random_std2d(30, 41, with_B=True, seed=5, batch=batch)
    land = c["F"] == cases.UNDEF
    c["F"] *= np.linspace(0.01, 100.0, batch)[:, None, None] ** 3          # very different loop counts
    c["F"][land] = cases.UNDEF
    p = c["p"]
    S = c["S0"].copy()
    fl, st = solvers.solve_standard_2D(S, c["A"], c["B"], c["C"], c["F"], "extend", "fixed", p["del1Sqr"], p["ratioQtr"],
                                       p["ratioSqr"], 1.2, cases.UNDEF, (0.0, 1.0, 0.0), 400, 1e-6, engine="resident")
    assert st["engine"] == "resident"
    loops = set()
    for b in range(0, batch, 7):
        cb = {k: (v[b] if isinstance(v, np.ndarray) and v.ndim == 3 else v) for k, v in c.items()}
        S_o, f_o = cases.run_std2d(oracle, cb, "extend", "fixed", 400, 1e-6, omega=1.2, ordering="colour")
        assert np.array_equal(S[b], S_o)
        _check_flags(fl[b], f_o)
        loops.add(int(f_o[2]))
    assert len(loops) > 3


def test_resident_overflow_and_zero_exit(gpu_ctx):
    c = cases.random_std2d(20, 30, with_B=False, seed=3, land=0.0)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 5000, 1e-12, omega=7.0, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 5000, 1e-12, omega=7.0, engine="resident")
    assert f_o[0] == 1.0 and f_g[0] == 1.0 and f_g[2] == f_o[2]
    assert np.array_equal(S_g, S_o, equal_nan=True)
    c["F"][:] = 0.0
    c["S0"][:] = 0.0                          # norm == 0 exit (numbas.py:410), 2-D standard form only
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 50, 1e-12, omega=1.4, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 50, 1e-12, omega=1.4, engine="resident")
    assert f_g[2] == f_o[2] == 0 and np.array_equal(S_g, S_o)


@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_resident_fourcolour_bridge_golden(gpu_ctx, bcx):
    """The 4-colour iteration carried out by the reference's own code (masked one-sweep calls)."""
    c, out = golden_io.load("bridge_std2d_9pt")
    S, _ = cases.run_std2d(xb, c, "fixed", bcx, 4, -1.0, omega=1.2, engine="resident")
    assert np.array_equal(S, out[f"S_{bcx}"])


@pytest.mark.parametrize("bcx", ["fixed", "periodic"])
def test_resident_redblack_bridge_goldens(gpu_ctx, bcx):
    c, out = golden_io.load("bridge_std2d")
    S, _ = cases.run_std2d(xb, c, "fixed", bcx, 5, -1.0, omega=1.4, engine="resident")
    assert np.array_equal(S, out[f"S_{bcx}"])
    c, out = golden_io.load("bridge_gen2d")
    S, _ = cases.run_gen2d(xb, c, "fixed", bcx, 5, -1.0, omega=1.3, engine="resident")
    assert np.array_equal(S, out[f"S_{bcx}"])
    c, out = golden_io.load("bridge_std2d_extend")
    S, _ = cases.run_std2d(xb, c, "extend", bcx, 5, -1.0, omega=1.4, engine="resident")
    assert np.array_equal(S, out[f"S_{bcx}"])


def test_auto_picks_resident_for_small_ninepoint(gpu_ctx):
    """engine='auto': a small slice the fused engine does not take (B != 0) runs resident, a large
    one still goes to the colour engine; too large for shared memory -> XinvError when forced."""
    c = cases.random_std2d(37, 73, with_B=True, seed=1)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 30, -1.0, omega=1.2, ordering="colour")
    S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 30, -1.0, omega=1.2)
    assert _engine(gpu_ctx) == "resident"
    assert np.array_equal(S_g, S_o)
    big = cases.random_std2d(130, 257, with_B=True, seed=2)
    cases.run_std2d(xb, big, "fixed", "fixed", 3, -1.0, omega=1.2)
    assert _engine(gpu_ctx) == "colour"
    with pytest.raises(xb.XinvError):
        cases.run_std2d(xb, big, "fixed", "fixed", 3, -1.0, omega=1.2, engine="resident")
