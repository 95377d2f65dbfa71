"""GPU: seeded random sweep over shapes, boundary conditions, land fractions, batch sizes, sweep
counts / tolerances and kernel families (fused general, fused RC, fused general-form, colour engine,
9-point, 3-D) against the ordering-matched C oracle -- bit-exact fields and identical loop counts.
Complements the structured cases of test_gpu_fused*.py / test_gpu_parity.py with combinations nobody
wrote down (odd sizes next to strip boundaries, strips of one or two rows, batches that freeze at
different sweeps, tolerances that fire mid-pass)."""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases

pytestmark = pytest.mark.gpu
BCY = ["fixed", "extend"]
BCX = ["fixed", "periodic"]


def _one_case(rng, k):
    fam = ["std_general", "std_rc", "gen_rc", "gen_general", "std_9pt", "std3d"][k % 6]
    ny, nx = int(rng.integers(3, 140)), int(rng.integers(4, 300))
    bcy, bcx = BCY[int(rng.integers(2))], BCX[int(rng.integers(2))]
    land = float(rng.choice([0.0, 0.05, 0.3, 0.9]))
    seed = int(rng.integers(1 << 30))
    if rng.random() < 0.5:
        mx, tol = int(rng.integers(0, 12)), -1.0
    else:
        mx, tol = 400, float(10.0 ** rng.uniform(-7, -2))
    omega = float(rng.choice([1.0, 1.4, 1.9]))
    return fam, ny, nx, bcy, bcx, land, seed, mx, tol, omega


@pytest.mark.parametrize("block", range(6))
def test_random_cases_bit_exact(gpu_ctx, monkeypatch, block):
    rng = np.random.default_rng(20261017 + block)
    monkeypatch.setenv("XINV_FUSED_RB", str(int(rng.choice([0, 0, 2, 6, 10, 24]))))     # 0 = automatic strip height
    for k in range(36):
        fam, ny, nx, bcy, bcx, land, seed, mx, tol, omega = _one_case(rng, k)
        tag = (fam, ny, nx, bcy, bcx, land, seed, mx, tol, omega)
        if fam == "std3d":
            nz, ny3, nx3 = int(rng.integers(3, 9)), min(ny, 40), min(nx, 60)
            c = cases.random_std3d(nz, ny3, nx3, seed, land=land)
            S_o, f_o = cases.run_std3d(oracle, c, bcy, bcx, mx, tol, omega=omega, ordering="colour")
            S_g, f_g = cases.run_std3d(xb, c, bcy, bcx, mx, tol, omega=omega)
        elif fam.startswith("gen"):
            c = (cases.random_gen2d_rowcoef(ny, nx, seed, land=land) if fam == "gen_rc"
                 else cases.random_gen2d(ny, nx, False, seed, land=land))
            S_o, f_o = cases.run_gen2d(oracle, c, bcy, bcx, mx, tol, omega=omega, ordering="colour")
            S_g, f_g = cases.run_gen2d(xb, c, bcy, bcx, mx, tol, omega=omega)
        else:
            c = (cases.random_std2d_rowcoef(ny, nx, seed, land=land) if fam == "std_rc"
                 else cases.random_std2d(ny, nx, fam == "std_9pt", seed, land=land))
            S_o, f_o = cases.run_std2d(oracle, c, bcy, bcx, mx, tol, omega=omega, ordering="colour")
            S_g, f_g = cases.run_std2d(xb, c, bcy, bcx, mx, tol, omega=omega)
        assert np.array_equal(S_g, S_o, equal_nan=True), (tag, xb.default_context().stats()["engine"])
        assert f_g[0] == f_o[0] and f_g[2] == f_o[2], (tag, f_g, f_o)


def test_random_batches_freeze_independently(gpu_ctx):
    """Batches of 2-5 slices with different forcing amplitudes: every slice equals its own oracle run."""
    rng = np.random.default_rng(7)
    for k in range(10):
        B, ny, nx = int(rng.integers(2, 6)), int(rng.integers(8, 80)), 2 * int(rng.integers(4, 90))
        bcy, bcx = BCY[int(rng.integers(2))], BCX[int(rng.integers(2))]
        c = cases.random_std2d_rowcoef(ny, nx, int(rng.integers(1 << 30)), batch=B) if k % 2 else \
            cases.random_std2d(ny, nx, False, int(rng.integers(1 << 30)), batch=B)
        if k % 2:                                       # row coefficients shared by the batch
            c["A"], c["C"] = c["A"], c["C"]
        for b in range(B):
            c["F"][b][c["F"][b] != cases.UNDEF] *= (1.0 + 4.0 * b)
        p = c["p"]
        tol = float(10.0 ** rng.uniform(-6, -3))
        S = c["S0"].copy()
        fl, st = xb.solve_standard_2D(S, c["A"], None, c["C"], c["F"], bcy, bcx, p["del1Sqr"], p["ratioQtr"],
                                      p["ratioSqr"], 1.4, mxLoop=600, tolerance=tol)
        for b in range(B):
            cb = dict(A=c["A"] if c["A"].ndim == 2 else c["A"][b], C=c["C"] if c["C"].ndim == 2 else c["C"][b],
                      F=c["F"][b], S0=c["S0"][b], p=p)
            S_o, f_o = cases.run_std2d(oracle, cb, bcy, bcx, 600, tol, omega=1.4, ordering="colour")
            assert fl[b, 2] == f_o[2] and np.array_equal(S[b], S_o), (k, b, st["engine"], fl[b], f_o)
