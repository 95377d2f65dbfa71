"""GPU parity at BASELINE.json's full sizes.

The C oracle does a 3600x1800 sweep in ~50 ms, so at full size it can still be the checker for
a handful of sweeps (bit-exact, ordering-matched); beyond that the engines are checked against
each other -- the RC kernels, the general kernels and the colour engine implement the same
iteration in three different ways and must agree to the bit after hundreds of sweeps -- and
through size-independent properties (linearity of the solve in the forcing)."""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases

pytestmark = pytest.mark.gpu


def _c2(seed=1000):
    return cases.poisson_latlon(1800, 3600, land=True, noise=1e-6, seed=seed)


def test_c2_full_size_vs_oracle_few_sweeps(gpu_ctx, monkeypatch):
    c = _c2()
    S_o, f_o = cases.run_std2d(oracle, c, "extend", "periodic", 6, -1.0, ordering="colour")        # 7 sweeps
    for rc in ("1", "0"):
        monkeypatch.setenv("XINV_FUSED_RC", rc)
        S_g, f_g = cases.run_std2d(xb, c, "extend", "periodic", 6, -1.0, engine="fused")
        assert xb.default_context().stats()["row_coeffs"] == int(rc)
        assert np.array_equal(S_g, S_o) and f_g[2] == f_o[2]
    S_c, f_c = cases.run_std2d(xb, c, "extend", "periodic", 6, -1.0, engine="colour")
    assert np.array_equal(S_c, S_o)


def test_c2_full_size_engines_agree_after_300_sweeps(gpu_ctx, monkeypatch):
    c = _c2(seed=7)
    out = {}
    for name, env, engine in (("rc", "1", "fused"), ("general", "0", "fused"), ("colour", "1", "colour")):
        monkeypatch.setenv("XINV_FUSED_RC", env)
        out[name] = cases.run_std2d(xb, c, "extend", "periodic", 300, -1.0, engine=engine)
    assert np.array_equal(out["rc"][0], out["general"][0])
    assert np.array_equal(out["rc"][0], out["colour"][0])
    assert out["rc"][1][2] == out["general"][1][2] == out["colour"][1][2] == 300
    assert np.abs(out["rc"][0]).max() > 0 and np.isfinite(out["rc"][0]).all()
    land = (c["F"] == cases.UNDEF)
    assert not out["rc"][0][land].any()                 # masked cells keep their initial value (numbas.py:344-348)


def test_c2_full_size_linearity(gpu_ctx):
    """SOR sweeps are linear in (F, S0): solve(a*F1 + F2) after n sweeps == a*solve(F1) + solve(F2)
    up to round-off (the property tests/test_Geopotential.py:92-104 of the reference checks)."""
    c1, c2 = _c2(seed=11), _c2(seed=12)
    ocean = c1["F"] != cases.UNDEF
    c12 = dict(c1, F=np.where(ocean, 2.5 * c1["F"] + c2["F"], cases.UNDEF))
    S1, _ = cases.run_std2d(xb, c1, "extend", "periodic", 99, -1.0)
    S2, _ = cases.run_std2d(xb, c2, "extend", "periodic", 99, -1.0)
    S12, _ = cases.run_std2d(xb, c12, "extend", "periodic", 99, -1.0)
    ref = 2.5 * S1 + S2
    assert np.abs(S12 - ref).max() <= 1e-10 * np.abs(ref).max()


def test_c5_slice_size_batch_vs_oracle(gpu_ctx):
    """Four 1440x720 slices over shared coefficients, 5 sweeps each, against the oracle."""
    B = 4
    c = cases.poisson_latlon(720, 1440, land=True, noise=1e-6, seed=5, batch=B)
    p = c["p"]
    S = c["S0"].copy()
    fl, st = xb.solve_standard_2D(S, c["A"], None, c["C"], c["F"], "fixed", "periodic", p["del1Sqr"], p["ratioQtr"],
                                  p["ratioSqr"], p["optArg"], mxLoop=4, tolerance=-1.0)
    assert st["engine"] == "fused" and st["row_coeffs"] == 1
    for b in range(B):
        cb = dict(A=c["A"], C=c["C"], F=c["F"][b], S0=c["S0"][b], p=p)
        S_o, f_o = cases.run_std2d(oracle, cb, "fixed", "periodic", 4, -1.0, ordering="colour")
        assert np.array_equal(S[b], S_o) and fl[b, 2] == f_o[2]


def test_c3_full_size_vs_oracle_few_sweeps(gpu_ctx):
    """invert_standard_3D at 37x180x360 (BASELINE configs[2]), 4 sweeps, bit-exact."""
    c = cases.random_std3d(37, 180, 360, seed=3)
    S_o, f_o = cases.run_std3d(oracle, c, "fixed", "periodic", 3, -1.0, ordering="colour")
    S_g, f_g = cases.run_std3d(xb, c, "fixed", "periodic", 3, -1.0)
    assert np.array_equal(S_g, S_o) and f_g[2] == f_o[2]


def test_c4_full_size_general_form_vs_oracle(gpu_ctx):
    """invert_general_2D at 360x720 (BASELINE configs[3] size), row-constant coefficients, 40 sweeps."""
    c = cases.random_gen2d_rowcoef(360, 720, seed=8)
    S_o, f_o = cases.run_gen2d(oracle, c, "fixed", "periodic", 39, -1.0, ordering="colour")
    S_g, f_g = cases.run_gen2d(xb, c, "fixed", "periodic", 39, -1.0)
    assert xb.default_context().stats()["engine"] == "fused"
    assert np.array_equal(S_g, S_o) and f_g[2] == f_o[2]
