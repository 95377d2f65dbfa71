"""GPU parity of the plane-marching fused engine for invert_standard_3D (xinv_march3d.cuh: one
red+black iteration, the y-extend rows, the norm and the loop control per pass) against the
ordering-matched C oracle: BIT-EXACT fields, identical loop counts.

XINV_FUSED3_VARIANT picks the tile height (X3_VARIANTS); the AROW kernels are chosen automatically when A is
constant along x (invert_omega's A = f^2 cos(lat)), the row-value kernels (X3_ROWS) when B and C are too
(invert_omega with N2 a profile along the levels, as in the reference's notebook 11); XINV_FUSED3_AROW=0 /
XINV_FUSED3_ROWS=0 force the more general kernels."""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases

pytestmark = pytest.mark.gpu

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]
SHAPES = [(3, 3, 4), (5, 9, 12), (4, 13, 60), (7, 12, 64), (6, 29, 61), (9, 40, 122), (5, 57, 130), (12, 25, 258)]
VARIANTS = ["0", "1", "2", "3", "4", "5"]      # xinv_march3d.cuh: X3_VARIANTS (tile height, ring depth, CTAs per SM)


def _arow(c, rows=False):
    """The same problem with A (rows: and B and C) constant along x (a different value per level and row).
    arow = False / True / "rows" -> stats row_coeffs 0 / 1 / 2."""
    out = dict(c)
    for k in ("A", "B", "C") if rows else ("A",):
        out[k] = np.ascontiguousarray(np.broadcast_to(c[k][..., :1], c[k].shape))
    return out


def _rc(arow):
    return 2 if arow == "rows" else int(bool(arow))


def _check(c, bcy, bcx, mx, tol=-1.0, omega=None, engine="fused", expect="fused", arow=None):
    S_o, f_o = cases.run_std3d(oracle, c, bcy, bcx, mx, tol, omega=omega, ordering="colour")
    S_g, f_g = cases.run_std3d(xb, c, bcy, bcx, mx, tol, omega=omega, engine=engine)
    st = xb.default_context().stats()
    assert st["engine"] == expect
    if arow is not None:
        assert st["row_coeffs"] == _rc(arow)
    assert np.array_equal(S_g, S_o), f"max diff {np.abs(S_g - S_o).max()} at {np.argwhere(S_g != S_o)[:5]}"
    assert f_g[0] == f_o[0] and f_g[2] == f_o[2]
    assert np.isclose(f_g[1], f_o[1], rtol=1e-6, atol=1e-13)    # a difference of two norms: tree sum vs serial sum
    return st


@pytest.mark.parametrize("arow", [False, True, "rows"])
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("bcy,bcx", BCS)
@pytest.mark.parametrize("shape", SHAPES)
def test_fused3d_bit_exact(gpu_ctx, monkeypatch, arow, variant, bcy, bcx, shape):
    monkeypatch.setenv("XINV_FUSED3_VARIANT", variant)
    if bcx == "periodic" and shape[2] % 2:
        pytest.skip("odd nx + periodic-x uses the wrap-fix colours (colour engine)")
    c = cases.random_std3d(*shape, seed=shape[0] * 10000 + shape[1] * 100 + shape[2])
    if arow:
        c = _arow(c, rows=(arow == "rows"))
    for mx in (0, 1, 4):
        _check(c, bcy, bcx, mx, arow=arow)


@pytest.mark.parametrize("arow", [False, True, "rows"])
@pytest.mark.parametrize("variant", ["0", "2", "3"])
@pytest.mark.parametrize("ntz", ["2", "3", "5"])
@pytest.mark.parametrize("bcy,bcx", BCS)
def test_fused3d_level_ranges_split_over_tiles(gpu_ctx, monkeypatch, arow, variant, ntz, bcy, bcx):
    """Small volumes split the levels over several tiles (2 halo levels each): every range boundary inside the grid."""
    monkeypatch.setenv("XINV_FUSED3_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED3_NTZ", ntz)
    for shape in ((11, 21, 64), (26, 14, 70)):
        c = cases.random_std3d(*shape, seed=shape[0] + int(ntz))
        if arow:
            c = _arow(c, rows=(arow == "rows"))
        for mx in (0, 3):
            _check(c, bcy, bcx, mx, arow=arow)


@pytest.mark.parametrize("arow", [False, True, "rows"])
@pytest.mark.parametrize("variant", ["6", "7", "8", "9", "10", "11"])
@pytest.mark.parametrize("ntz", ["1", "2"])
def test_fused3d_deep_rings_and_tall_tiles(gpu_ctx, monkeypatch, arow, variant, ntz):
    """The variants the automatic choice uses besides 0 (deeper rings, 20- and 24-row tiles), whole and split in two level ranges."""
    monkeypatch.setenv("XINV_FUSED3_VARIANT", variant)
    monkeypatch.setenv("XINV_FUSED3_NTZ", ntz)
    for shape, (bcy, bcx) in (((9, 40, 122), ("extend", "periodic")), ((21, 45, 70), ("fixed", "fixed"))):
        c = cases.random_std3d(*shape, seed=shape[1] + int(variant))
        if arow:
            c = _arow(c, rows=(arow == "rows"))
        _check(c, bcy, bcx, 3, arow=arow)


def test_fused3d_arow_equals_general_kernels(gpu_ctx, monkeypatch):
    """The same x-constant A through the general kernels (XINV_FUSED3_AROW=0)."""
    c = _arow(cases.random_std3d(9, 40, 122, seed=77))
    c["A"][3, 7, :] = cases.UNDEF              # a whole row of undef A
    _check(c, "extend", "periodic", 5, arow=True)
    monkeypatch.setenv("XINV_FUSED3_AROW", "0")
    _check(c, "extend", "periodic", 5, arow=False)


def test_fused3d_rows_equals_general_kernels(gpu_ctx, monkeypatch):
    """x-constant A, B and C (whole rows of undef among them) through the row-value kernels, the AROW kernels
    (XINV_FUSED3_ROWS=0) and the general kernels (XINV_FUSED3_AROW=0)."""
    c = _arow(cases.random_std3d(9, 40, 122, seed=78), rows=True)
    c["B"][4, 9, :] = cases.UNDEF
    c["C"][2, 30, :] = cases.UNDEF
    c["A"][6, 1, :] = cases.UNDEF
    _check(c, "extend", "periodic", 5, arow="rows")
    _check(c, "fixed", "fixed", 5, arow="rows")
    monkeypatch.setenv("XINV_FUSED3_ROWS", "0")
    _check(c, "extend", "periodic", 5, arow=True)
    monkeypatch.setenv("XINV_FUSED3_AROW", "0")
    _check(c, "extend", "periodic", 5, arow=False)


def test_fused3d_rows_refused_when_one_value_differs(gpu_ctx):
    """One cell of C that differs from its row: the AROW kernels, not the row-value kernels."""
    c = _arow(cases.random_std3d(6, 20, 64, seed=79), rows=True)
    c["C"][3, 11, 40] *= 1.0 + 2.0 ** -50
    _check(c, "fixed", "periodic", 4, arow=True)
    c["A"][2, 5, 63] *= 1.0 + 2.0 ** -50
    _check(c, "fixed", "periodic", 4, arow=False)


@pytest.mark.parametrize("arow", [False, "rows"])
@pytest.mark.parametrize("bcy,bcx", BCS)
def test_fused3d_auto_variant_and_tolerance(gpu_ctx, bcy, bcx, arow):
    """Default variant choice, solved to tolerance: identical loop counts."""
    c = cases.random_std3d(10, 44, 72, seed=11, land=0.05)
    if arow:
        c = _arow(c, rows=(arow == "rows"))
    st = _check(c, bcy, bcx, 400, tol=1e-7, engine="auto", arow=arow)
    assert st["sweeps_launched"] >= 1


def test_fused3d_odd_nx_periodic_falls_back_to_colour_engine(gpu_ctx):
    c = cases.random_std3d(6, 20, 31, seed=5)
    _check(c, "fixed", "periodic", 5, engine="auto", expect="colour")
    with pytest.raises(xb.XinvError) as ei:
        cases.run_std3d(xb, c, "fixed", "periodic", 5, -1.0, engine="fused")
    assert ei.value.code == -5


def test_fused3d_undef_psi_and_undef_coefficients(gpu_ctx):
    """undef values inside psi (not counted by the norm, not copied by y-extend) and inside A/B/C."""
    c = cases.random_std3d(8, 30, 64, seed=21)
    rng = np.random.default_rng(3)
    for k in ("A", "B", "C"):
        c[k][rng.random(c[k].shape) < 0.03] = cases.UNDEF
    c["S0"][rng.random(c["S0"].shape) < 0.05] = cases.UNDEF
    c["S0"][:, 1, 5:9] = cases.UNDEF          # rows that y-extend must not copy
    c["S0"][:, -2, 20:30] = cases.UNDEF
    for bcy, bcx in BCS:
        _check(c, bcy, bcx, 3)


@pytest.mark.parametrize("arow", [True, "rows"])
def test_fused3d_batched_shared_coefficients_and_freezing(gpu_ctx, arow):
    """A batch of volumes sharing A, B, C (stride 0), each stopping on its own test."""
    nb, shape = 5, (7, 26, 64)
    c = _arow(cases.random_std3d(*shape, seed=8, land=0.05), rows=(arow == "rows"))
    rng = np.random.default_rng(9)
    F = np.stack([c["F"] * (1.0 + 3.0 * t) for t in range(nb)])
    F[:, c["F"] == cases.UNDEF] = cases.UNDEF
    S0 = rng.standard_normal((nb,) + shape) * np.array([1.0, 1e-3, 10.0, 1e-6, 1.0])[:, None, None, None]
    p = c["p"]
    S = S0.copy()
    fl, st = xb.solve_standard_3D(S, c["A"], c["B"], c["C"], F, "fixed", "extend", "periodic", p["del1Sqr"],
                                  p["ratio2Sqr"], p["ratio1Sqr"], p["optArg"], cases.UNDEF, mxLoop=300, tolerance=1e-6)
    assert st["engine"] == "fused" and st["row_coeffs"] == _rc(arow)
    loops = set()
    for t in range(nb):
        ct = dict(c, F=np.ascontiguousarray(F[t]), S0=S0[t])
        S_o, f_o = cases.run_std3d(oracle, ct, "extend", "periodic", 300, 1e-6, ordering="colour")
        assert np.array_equal(S[t], S_o), t
        assert fl[t, 2] == f_o[2] and fl[t, 0] == f_o[0]
        loops.add(int(f_o[2]))
    assert len(loops) > 1, "slices were meant to stop at different iterations"


@pytest.mark.parametrize("arow", [False, True, "rows"])
def test_fused3d_batched_dense_coefficients(gpu_ctx, arow):
    """Every volume of the batch with its own A, B, C (dense, or one value per row)."""
    nb, shape = 3, (5, 20, 70)
    c = cases.random_std3d(*shape, seed=31, batch=nb)
    if arow:
        c = _arow(c, rows=(arow == "rows"))
    p = c["p"]
    S = c["S0"].copy()
    fl, st = xb.solve_standard_3D(S, c["A"], c["B"], c["C"], c["F"], "fixed", "fixed", "fixed", p["del1Sqr"],
                                  p["ratio2Sqr"], p["ratio1Sqr"], p["optArg"], cases.UNDEF, mxLoop=6, tolerance=-1.0)
    assert st["engine"] == "fused" and st["row_coeffs"] == _rc(arow)
    for t in range(nb):
        ct = dict(A=c["A"][t], B=c["B"][t], C=c["C"][t], F=c["F"][t], S0=c["S0"][t], p=p)
        S_o, f_o = cases.run_std3d(oracle, ct, "fixed", "fixed", 6, -1.0, ordering="colour")
        assert np.array_equal(S[t], S_o), t


def test_fused3d_warm_start_equals_one_long_solve(gpu_ctx):
    """S is in/out: two solves of 3 sweeps continue where one of 6 would be (animate_iteration's contract)."""
    c = cases.random_std3d(6, 30, 64, seed=14)
    S6, _ = cases.run_std3d(xb, c, "extend", "periodic", 5, -1.0, engine="fused")
    S3, _ = cases.run_std3d(xb, c, "extend", "periodic", 2, -1.0, engine="fused")
    c2 = dict(c, S0=S3)
    S33, _ = cases.run_std3d(xb, c2, "extend", "periodic", 2, -1.0, engine="fused")
    assert np.array_equal(S6, S33)


def test_fused3d_overflow_flag(gpu_ctx):
    """A diverging iteration (omega far above 2) sets flags[0] exactly when the oracle does."""
    c = cases.random_std3d(6, 20, 64, seed=17, land=0.0)
    _check(c, "fixed", "periodic", 2000, tol=1e-30, omega=2.9)


def test_fused3d_equals_colour_engine_at_c3_size(gpu_ctx):
    """37 x 180 x 360 (BASELINE configs[2]): fused and colour engines give the same bits."""
    c = cases.random_std3d(37, 180, 360, seed=3, land=0.1)
    S_f, f_f = cases.run_std3d(xb, c, "extend", "periodic", 9, -1.0, engine="fused")
    S_c, f_c = cases.run_std3d(xb, c, "extend", "periodic", 9, -1.0, engine="colour")
    assert np.array_equal(S_f, S_c)
    assert f_f[2] == f_c[2]
