"""GPU: the xarray-in / xarray-out facade end to end through the C-ABI.

* reference known answers (tests/test_GillMatsuno.py:55-57 of the reference; SURVEY.md
  section 4 probes) with ``ordering='lexicographic'``: the reference's own trajectory,
  so loop counts and values must match to 1e-12 relative;
* the default red-black ordering against the same goldens at the reference tests' own
  tolerance (np.isclose default, rtol 1e-5) and against the ordering-matched oracle
  (bit-exact) on the BASELINE-shaped problems C1, C3, C4.
"""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from xinvert_b200 import core
from tests import oracle_backend
from tests.test_apps_host import _c1_zeta, _grid2

pytestmark = pytest.mark.gpu
DA = xb.DataArray


def _gm_inputs():
    lon, lat = np.linspace(0, 360, 144), np.linspace(-90, 90, 73)
    la, lo, coords = _grid2(73, 144, lat, lon)
    Q1 = 0.05 * np.exp(-((la - 0) ** 2 + (lo - 120) ** 2) / 100.0)
    Q2 = 0.05 * np.exp(-((la - 10) ** 2 + (lo - 120) ** 2) / 100.0) - 0.05 * np.exp(-((la + 10) ** 2 + (lo - 120) ** 2) / 100.0)
    Q3 = 0.05 * np.exp(-((la - 10) ** 2 + (lo - 120) ** 2) / 100.0)
    return np.stack([Q1, Q2, Q3]), coords


def test_gill_matsuno_reference_golden_lexicographic(gpu_ctx, capsys):
    """All three heatings as ONE batched call (a 'case' dim), reference ordering."""
    Q, coords = _gm_inputs()
    iParams = {'BCs': ['fixed', 'periodic'], 'mxLoop': 2000, 'tolerance': 1e-8, 'optArg': 1.4,
               'ordering': 'lexicographic'}
    mParams = {'epsilon': 1e-5, 'Phi': 5000}
    h = xb.invert_GillMatsuno(DA(Q, ['case', 'lat', 'lon'], dict(coords, case=np.arange(3))), dims=['lat', 'lon'],
                              iParams=iParams, mParams=mParams)
    out = capsys.readouterr().out
    for loops in (1628, 1146, 1618):
        assert f"loops {loops:4.0f} and tolerance is" in out
    u, v = xb.cal_flow(h, dims=['lat', 'lon'], BCs=['fixed', 'periodic'], mParams=mParams, vtype='GillMatsuno')
    ke = ((u.values ** 2 + v.values ** 2) / 2).sum(axis=(1, 2))
    assert np.allclose(ke, [4351.62244687, 5833.33192343, 5100.85325027], rtol=1e-10, atol=0)


def test_gill_matsuno_reference_golden_default_ordering(gpu_ctx):
    """Default (red-black) ordering: the reference test's own assertions hold."""
    Q, coords = _gm_inputs()
    iParams = {'BCs': ['fixed', 'periodic'], 'mxLoop': 2000, 'tolerance': 1e-8, 'optArg': 1.4, 'printInfo': False}
    mParams = {'epsilon': 1e-5, 'Phi': 5000}
    h = xb.invert_GillMatsuno(DA(Q, ['case', 'lat', 'lon'], dict(coords, case=np.arange(3))), dims=['lat', 'lon'],
                              iParams=iParams, mParams=mParams)
    u, v = xb.cal_flow(h, dims=['lat', 'lon'], BCs=['fixed', 'periodic'], mParams=mParams, vtype='GillMatsuno')
    ke = ((u.values ** 2 + v.values ** 2) / 2).sum(axis=(1, 2))
    assert (h.values[0] <= 0).all() and (np.abs(h.values[1]) <= 370).all() and (h.values[2] <= 0).all()
    assert np.isclose(ke[0], 4351.62244687) and np.isclose(ke[1], 5833.33192343) and np.isclose(ke[2], 5100.85325027)


def test_poisson_c1_reference_known_answer(gpu_ctx, capsys):
    """BASELINE configs[0]: lexicographic order reproduces the reference run exactly
    (loop 2380, max|psi| 13182413.993245527); the default order converges to the same
    field within the iteration tolerance and in a comparable number of sweeps."""
    zeta, coords = _c1_zeta()
    ip = {'BCs': ['fixed', 'periodic'], 'optArg': 1.4, 'tolerance': 1e-8, 'mxLoop': 5000}
    F = DA(zeta, ['lat', 'lon'], coords)
    lex = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=dict(ip, ordering='lexicographic'))
    assert "loops 2380" in capsys.readouterr().out
    assert np.isclose(np.abs(lex.values).max(), 13182413.993245527, rtol=1e-12)
    rb = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ip)
    loops = int(capsys.readouterr().out.split("loops")[1].split()[0])
    assert 1500 < loops < 3500
    assert np.abs(rb.values - lex.values).max() / np.abs(lex.values).max() < 1e-4


def _via_oracle(fn, *a, **k):
    """Same facade call with the C oracle as the backend (host logic identical)."""
    saved = core.solvers
    core.solvers = oracle_backend
    try:
        return fn(*a, **k)
    finally:
        core.solvers = saved


def test_c3_omega_3d_bit_exact_vs_oracle(gpu_ctx):
    """BASELINE configs[2] shape at reduced size in y/x (the oracle must finish in seconds):
    37 levels, variable N2(lev,lat,lon), fixed/fixed/periodic."""
    nz, ny, nx = 37, 60, 120
    lev = 100000.0 - 2500.0 * np.arange(nz)
    lat = -88.5 + 3.0 * np.arange(ny)
    lon = 3.0 * np.arange(nx)
    rng = np.random.default_rng(1)
    coords = {'LEV': lev, 'lat': lat, 'lon': lon}
    N2 = DA(1e-6 * (1 + 0.5 * rng.random((nz, ny, nx))), ['LEV', 'lat', 'lon'], coords)
    F = DA(1e-17 * rng.standard_normal((nz, ny, nx)), ['LEV', 'lat', 'lon'], coords)
    ip = {'BCs': ['fixed', 'fixed', 'periodic'], 'tolerance': 1e-10, 'mxLoop': 60, 'printInfo': False}
    kw = dict(dims=['LEV', 'lat', 'lon'], iParams=ip, mParams={'N2': N2})
    w_g = xb.invert_omega(F, **kw)
    w_o = _via_oracle(xb.invert_omega, F, **kw)
    assert np.array_equal(w_g.values, w_o.values)


def test_c4_gill_matsuno_beta_plane_bit_exact_vs_oracle(gpu_ctx):
    """BASELINE configs[3]: 720x360 cartesian beta-plane (SURVEY.md 8d recipe), 150 sweeps."""
    ny, nx = 360, 720
    y = np.linspace(-5e6, 5e6, ny)
    x = np.linspace(0, 4e7, nx, endpoint=False)
    yy, xx, coords = _grid2(ny, nx, y, x, 'y', 'x')
    Q = DA(0.05 * np.exp(-((yy / 1e6) ** 2 + ((xx - 2e7) / 2e6) ** 2)), ['y', 'x'], coords)
    ip = {'BCs': ['fixed', 'periodic'], 'optArg': 1.4, 'tolerance': -1.0, 'mxLoop': 149, 'printInfo': False}
    mp = {'f0': 0.0, 'beta': 2e-11, 'epsilon': 1e-5, 'Phi': 5000}
    kw = dict(dims=['y', 'x'], coords='cartesian', iParams=ip, mParams=mp)
    h_g = xb.invert_GillMatsuno(Q, **kw)
    h_o = _via_oracle(xb.invert_GillMatsuno, Q, **kw)
    assert np.array_equal(h_g.values, h_o.values) and np.abs(h_g.values).max() > 0
    u, v = xb.cal_flow(h_g, dims=['y', 'x'], coords='cartesian', mParams=mp, vtype='GillMatsuno')
    assert np.isfinite(u.values).all() and np.isfinite(v.values).all()


def test_c2_style_ocean_poisson_batched(gpu_ctx, capsys):
    """Helmholtz_ocean-style (tests/test_Poisson.py:44-65 of the reference): land mask,
    extend/periodic, several time slices; fused engine; each slice equals the oracle."""
    ny, nx, T = 90, 180, 4
    zeta, coords = _c1_zeta(ny, nx)
    rng = np.random.default_rng(0)
    z = np.stack([zeta * (1 + 0.5 * t) + 1e-6 * rng.standard_normal((ny, nx)) for t in range(T)])
    lam, phi = np.deg2rad(coords['lon'])[None, :], np.deg2rad(coords['lat'])[:, None]
    z[:, np.sin(5 * lam) * np.cos(3 * phi) > 0.6] = np.nan
    F = DA(z, ['time', 'lat', 'lon'], dict(coords, time=np.arange(T)))
    ip = {'BCs': ['extend', 'periodic'], 'tolerance': 1e-9, 'mxLoop': 5000}
    s_g = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ip)
    out_g = capsys.readouterr().out
    assert xb.default_context().stats()["engine"] == "cluster"      # four 90 x 180 slices: one thread-block cluster each
    s_o = _via_oracle(xb.invert_Poisson, F, dims=['lat', 'lon'], iParams=ip)
    out_o = capsys.readouterr().out
    assert np.array_equal(s_g.values, s_o.values, equal_nan=True)
    loops = lambda txt: [ln.split("loops")[1].split()[0] for ln in txt.strip().splitlines()]
    assert loops(out_g) == loops(out_o)


def test_linearity_property(gpu_ctx):
    """tests/test_Geopotential.py:92-104 of the reference, on synthetic data:
    invert(F1 + F2) == invert(F1) + invert(F2) to 5e-5 relative."""
    zeta, coords = _c1_zeta(60, 120)
    rng = np.random.default_rng(4)
    F1 = zeta
    F2 = 1e-5 * rng.standard_normal(zeta.shape) * np.cos(np.deg2rad(coords['lat']))[:, None]
    ip = {'BCs': ['fixed', 'periodic'], 'tolerance': 1e-12, 'mxLoop': 8000, 'printInfo': False}
    inv = lambda f: xb.invert_Poisson(DA(f, ['lat', 'lon'], coords), dims=['lat', 'lon'], iParams=ip).values
    s12, s1, s2 = inv(F1 + F2), inv(F1), inv(F2)
    assert np.allclose(s12, s1 + s2, rtol=5e-5, atol=5e-5 * np.abs(s12).max())


def test_stommel_ishida_reference_known_answers(gpu_ctx, capsys):
    """tests/test_Ishida.py:13-63 of the reference (general form, land strips, user undef -9999):
    with ordering='lexicographic' the GPU reproduces the reference's loop counts and maxima
    (SURVEY.md section 4 probe values); the default colour ordering (nx = 251 is odd and x is periodic:
    wrap-fix colours on the colour engine) is bit-equal to the ordering-matched oracle and meets the
    reference test's own bounds."""
    from tests.test_apps_host import _ishida_case
    curl, coords, _ = _ishida_case()
    curl[65:, 100:104] = -9999
    curl[:75, 130:134] = -9999
    F = DA(curl, ['ydef', 'xdef'], coords)
    land = (curl == -9999)
    base = {'BCs': ['fixed', 'periodic'], 'mxLoop': 3000, 'tolerance': 1e-9, 'optArg': 1.4, 'undef': -9999}
    want = {0.0009: ("loops 1474", 451695.81539746444, 5.5e5), 0.018: ("loops 3000", 26968.645791689938, 2.8e4)}
    for R, (loops, amax, bound) in want.items():
        mp = {'beta': 2.2e-11, 'R': R, 'D': 200}
        kw = dict(dims=['ydef', 'xdef'], coords='cartesian', mParams=mp)
        h = xb.invert_Stommel(F, iParams=dict(base, ordering='lexicographic'), **kw)
        assert loops in capsys.readouterr().out
        assert np.isclose(np.abs(h.values[~land]).max(), amax, rtol=1e-12)
        h_g = xb.invert_Stommel(F, iParams=dict(base, printInfo=False), **kw)
        st = xb.default_context().stats()
        assert st["engine"] == "colour" and st["ncolours"] == 4
        h_o = _via_oracle(xb.invert_Stommel, F, iParams=dict(base, printInfo=False), **kw)
        assert np.array_equal(h_g.values, h_o.values)
        assert (h_g.values[land] == -9999).all() and np.abs(h_g.values[~land]).max() <= bound


def test_eliassen_nine_point_bit_exact_vs_oracle(gpu_ctx):
    """invert_Eliassen (B != 0: 9-point stencil, 4-colour ordering; a z-r section of this size runs
    on the resident engine -- the whole solve in one CTA -- a larger one on the colour engine)."""
    ny, nx = 60, 90
    z, y = np.linspace(1000., 100., ny), np.linspace(0., 5e5, nx)
    coords = {'z': z, 'r': y}
    rng = np.random.default_rng(5)
    A = DA(1 + 0.2 * rng.random((ny, nx)), ['z', 'r'], coords)
    B = DA(0.1 * rng.standard_normal((ny, nx)), ['z', 'r'], coords)
    C = DA(1 + 0.2 * rng.random((ny, nx)), ['z', 'r'], coords)
    F = DA(1e-9 * rng.standard_normal((ny, nx)), ['z', 'r'], coords)
    ip = {'BCs': ['fixed', 'fixed'], 'mxLoop': 600, 'tolerance': 1e-10, 'optArg': 1.14, 'printInfo': False}
    kw = dict(dims=['z', 'r'], coords='cartesian', iParams=ip, mParams={'A': A, 'B': B, 'C': C})
    s_g = xb.invert_Eliassen(F, **kw)
    st = xb.default_context().stats()
    assert st["engine"] == "resident" and st["ncolours"] == 4
    s_o = _via_oracle(xb.invert_Eliassen, F, **kw)
    assert np.array_equal(s_g.values, s_o.values)


@pytest.mark.parametrize("coef_dims", ["core", "full", "mixed"])
def test_eliassen_device_front_end_equals_host_path(gpu_ctx, monkeypatch, coef_dims):
    """invert_Eliassen with icbc=None goes through xinv_std2d_front (masks, zero guess, de-masking on the device; the
    coefficient fields handed over as the user holds them, core-shaped ones shared by the batch); the reference-shaped
    host path must give the same bits: NaN-marked land, a batch over time, every form of A / B / C."""
    from xinvert_b200 import apps
    T, ny, nx = 5, 37, 73
    z, y = np.linspace(100000., 10000., ny), np.linspace(-80., 80., nx)
    co = {'time': np.arange(T), 'lev': z, 'lat': y}
    rng = np.random.default_rng(8)
    core_co = {'lev': z, 'lat': y}
    mk = lambda v, full: DA(v, ['time', 'lev', 'lat'], co) if full else DA(v, ['lev', 'lat'], core_co)
    full = {"core": (False, False, False), "full": (True, True, True), "mixed": (True, False, False)}[coef_dims]
    shp = lambda f: (T, ny, nx) if f else (ny, nx)
    A = mk(1 + 0.2 * rng.random(shp(full[0])), full[0])
    B = mk(0.1 * rng.standard_normal(shp(full[1])), full[1])
    C = DA(1 + 0.2 * rng.random(ny), ['lev'], {'lev': z}) if coef_dims == "mixed" else mk(1 + 0.2 * rng.random(shp(full[2])), full[2])
    Fv = 1e-9 * rng.standard_normal((T, ny, nx))
    Fv[:, 30:, 20:30] = np.nan                                     # topography
    F = DA(Fv, ['time', 'lev', 'lat'], co)
    ip = {'BCs': ['fixed', 'fixed'], 'mxLoop': 300, 'tolerance': 1e-9, 'optArg': 1.2, 'printInfo': False}
    kw = dict(dims=['lev', 'lat'], coords='z-lat', mParams={'A': A, 'B': B, 'C': C})
    calls = []
    real = apps._device_solvers.solve_standard_2D_front
    monkeypatch.setattr(apps._device_solvers, "solve_standard_2D_front", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    ip_f = dict(ip)
    s_f = xb.invert_Eliassen(F, iParams=ip_f, **kw)
    assert calls, "the device front end was not used"
    assert gpu_ctx.stats()["engine"] == "resident" and ip_f['stats']['h2d_bytes'] < 8 * (Fv.size * (1 + sum(full)) + 3 * ny * nx) + 4096
    monkeypatch.setattr(apps, "_eliassen_device_front", lambda *a, **k: None)
    ip_h = dict(ip)
    s_h = xb.invert_Eliassen(F, iParams=ip_h, **kw)
    assert np.array_equal(s_f.values, s_h.values, equal_nan=True)
    assert np.array_equal(ip_f['flags_all'][:, [0, 2]], ip_h['flags_all'][:, [0, 2]])
    assert np.isnan(s_f.values[:, 30:, 20:30]).all() and np.isfinite(s_f.values[:, :30]).all()
    # an unmasked non-finite forcing value: falls back to the host path (which reproduces the reference's NaN smear)
    Fv2 = Fv.copy(); Fv2[0, 5, 5] = np.inf
    monkeypatch.undo()
    s_bad = xb.invert_Eliassen(DA(Fv2, ['time', 'lev', 'lat'], co), iParams=dict(ip), **kw)
    assert s_bad.values.shape == Fv.shape


def test_stommel_idealized_fused_general_form(gpu_ctx, capsys):
    """tests/test_StommelWBC.py:14-45 of the reference (general_2D halves): lexicographic ordering
    reproduces the reference's loop counts / maxima; the default ordering runs on the fused
    general-form kernels (fixed/fixed, coefficients constant along x) bit-equal to the oracle."""
    xnum, ynum = 201, 151
    Lx, Ly = 1e7, 2 * np.pi * 1e6
    x, y = np.linspace(0, Lx, xnum), np.linspace(0, Ly, ynum)
    yg, xg, coords = _grid2(ynum, xnum, y, x, 'ydef', 'xdef')
    curl = DA(-0.3 * np.sin(np.pi * yg / Ly) * np.pi / Ly, ['ydef', 'xdef'], coords)
    base = {'BCs': ['fixed', 'fixed'], 'mxLoop': 5000, 'optArg': 1.9, 'tolerance': 1e-12}
    for beta, loops, mx in ((0, 3213, 611203.653077336), (1.8e-11, 457, 282080.3876195683)):
        kw = dict(dims=['ydef', 'xdef'], coords='cartesian', mParams={'beta': beta, 'R': 0.0008, 'D': 200})
        S = xb.invert_Stommel(curl, iParams=dict(base, ordering='lexicographic'), **kw)
        assert f"loops {loops:4.0f}" in capsys.readouterr().out
        assert np.isclose(S.max(), mx, rtol=1e-12)
        S_g = xb.invert_Stommel(curl, iParams=dict(base, printInfo=False), **kw)
        st = xb.default_context().stats()
        assert st["engine"] in ("fused", "cluster") and st["row_coeffs"] == 1
        S_o = _via_oracle(xb.invert_Stommel, curl, iParams=dict(base, printInfo=False), **kw)
        assert np.array_equal(S_g.values, S_o.values)
        assert np.isclose(S_g.max(), mx, rtol=1e-6)          # same fixed point as the reference ordering


@pytest.mark.parametrize("coords,bcs", [("lat-lon", ["extend", "periodic"]), ("lat-lon", ["fixed", "periodic"]),
                                        ("cartesian", ["fixed", "fixed"])])
def test_poisson_device_front_end_equals_host_path(gpu_ctx, monkeypatch, coords, bcs):
    """invert_Poisson with icbc=None goes through xinv_std2d_rows (masking, coefficients, forcing
    scale and de-masking on the device); forcing it onto the reference-shaped host path must give the
    same bits, the same printed loop counts and the same undef pattern -- NaN-marked and
    value-marked land, batched over a time axis."""
    from xinvert_b200 import apps
    ny, nx, T = 90, 180, 3
    zeta, co = _c1_zeta(ny, nx)
    rng = np.random.default_rng(2)
    z = np.stack([zeta * (1 + 0.5 * t) + 1e-6 * rng.standard_normal((ny, nx)) for t in range(T)])
    lam, phi = np.deg2rad(co['lon'])[None, :], np.deg2rad(co['lat'])[:, None]
    land = np.sin(5 * lam) * np.cos(3 * phi) > 0.6
    if coords == "cartesian":
        co = {'lat': 1e5 * np.arange(ny), 'lon': 1e5 * np.arange(nx)}
    for undef in (np.nan, -9999.0):
        zz = z.copy()
        zz[:, land] = undef
        F = DA(zz, ['time', 'lat', 'lon'], dict(co, time=np.arange(T)))
        ip = {'BCs': bcs, 'tolerance': 1e-9, 'mxLoop': 3000, 'undef': undef, 'printInfo': False}
        calls = []
        real = apps._device_solvers.solve_standard_2D_rows
        monkeypatch.setattr(apps._device_solvers, "solve_standard_2D_rows", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
        s_f = xb.invert_Poisson(F, dims=['lat', 'lon'], coords=coords, iParams=dict(ip))
        assert calls, "the device front end was not used"
        monkeypatch.setattr(apps, "_poisson_device_front", lambda *a, **k: None)
        s_h = xb.invert_Poisson(F, dims=['lat', 'lon'], coords=coords, iParams=dict(ip))
        monkeypatch.undo()
        assert np.array_equal(s_f.values, s_h.values, equal_nan=True)
        lv = s_f.values[:, land]
        assert np.isnan(lv).all() if np.isnan(undef) else (lv == undef).all()
        assert np.isfinite(s_f.values[:, ~land]).all() and np.abs(s_f.values[:, ~land]).max() > 0


def test_poisson_device_front_end_falls_back(gpu_ctx):
    """Inputs the front end does not take (odd nx with periodic-x; an infinite forcing value) still
    give the host path's answer."""
    zeta, co = _c1_zeta(30, 61)
    F = DA(zeta, ['lat', 'lon'], co)
    ip = {'BCs': ['fixed', 'periodic'], 'tolerance': 1e-9, 'mxLoop': 200, 'printInfo': False}
    s = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=dict(ip))
    s_o = _via_oracle(xb.invert_Poisson, F, dims=['lat', 'lon'], iParams=dict(ip))
    assert np.array_equal(s.values, s_o.values)
    zeta, co = _c1_zeta(30, 60)
    zeta[7, 9] = np.inf
    F = DA(zeta, ['lat', 'lon'], co)
    s = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=dict(ip))
    s_o = _via_oracle(xb.invert_Poisson, F, dims=['lat', 'lon'], iParams=dict(ip))
    assert np.array_equal(s.values, s_o.values, equal_nan=True)


def test_front_end_with_device_tensors(gpu_ctx):
    """solve_standard_2D_rows with the forcing already on the device (CUDA tensor in, CUDA tensor out)."""
    import torch
    ny, nx = 64, 128
    zeta, lat, lon = cases_user(ny, nx)
    lats = np.deg2rad(lat)
    cosG = np.cos(lats)
    latm = np.empty(ny); latm[0] = np.nan; latm[1:] = lats[:-1]
    A_rows, C_rows = np.cos((lats + latm) / 2.0), 1.0 / cosG
    from tests import cases
    p = cases.params2d(ny, nx, np.deg2rad(180.0 / ny) * cases.REARTH, np.deg2rad(360.0 / nx) * cases.REARTH)
    args = (A_rows, C_rows, cosG, np.nan, np.nan, "extend", "periodic", p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], 1.4)
    S_h, fl_h, _ = xb.solve_standard_2D_rows(zeta, *args, mxLoop=300, tolerance=1e-8)
    Fd = torch.from_numpy(zeta).to("cuda:0")
    S_d, fl_d, st = xb.solve_standard_2D_rows(Fd, *args, mxLoop=300, tolerance=1e-8)
    assert st["h2d_bytes"] == 0 and st["d2h_bytes"] == 0
    assert np.array_equal(S_d.cpu().numpy(), S_h, equal_nan=True) and np.array_equal(fl_d, fl_h)
    assert np.isnan(S_h[np.isnan(zeta)]).all() and np.isfinite(S_h[~np.isnan(zeta)]).all()


def cases_user(ny, nx):
    from tests import cases
    return cases.poisson_latlon_user(ny, nx, land=True, noise=1e-6, seed=3)


@pytest.mark.parametrize("model", ["GillMatsuno-latlon", "GillMatsuno-cartesian", "Stommel-cartesian", "Stommel-latlon"])
def test_general_form_device_front_end_equals_host_path(gpu_ctx, monkeypatch, model):
    """invert_GillMatsuno / invert_Stommel with icbc=None go through xinv_gen2d_rows; the host path must
    give the same bits (land marked by NaN and by a value, batched over a time axis)."""
    from xinvert_b200 import apps
    ny, nx, T = 60, 120, 2
    rng = np.random.default_rng(9)
    if model.endswith("latlon"):
        co = {'lat': np.linspace(-59, 59, ny), 'lon': np.linspace(0, 357, nx)}
        coords = 'lat-lon'
    else:
        co = {'lat': np.linspace(-3e6, 3e6, ny), 'lon': np.linspace(0, 2e7, nx, endpoint=False)}
        coords = 'cartesian'
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    q = np.stack([(1 + t) * 0.05 * np.exp(-((yy - 30) ** 2 + (xx - 60) ** 2) / 50.0) + 1e-4 * rng.standard_normal((ny, nx))
                  for t in range(T)])
    land = (np.abs(yy - 20) < 3) & (np.abs(xx - 30) < 10)
    if model.startswith("Gill"):
        fn, mp = xb.invert_GillMatsuno, {'epsilon': 1e-5, 'Phi': 5000, 'f0': 0.0, 'beta': 2e-11}
    else:
        fn, mp = xb.invert_Stommel, {'beta': 2e-11, 'R': 5e-4, 'D': 200}
        q = q * 1e-9
    for undef in (np.nan, -9999.0):
        qq = q.copy()
        qq[:, land] = undef
        F = DA(qq, ['time', 'lat', 'lon'], dict(co, time=np.arange(T)))
        ip = {'BCs': ['fixed', 'periodic'], 'tolerance': 1e-9, 'mxLoop': 400, 'optArg': 1.4, 'undef': undef, 'printInfo': False}
        calls = []
        real = apps._device_solvers.solve_general_2D_rows
        monkeypatch.setattr(apps._device_solvers, "solve_general_2D_rows", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
        s_f = fn(F, dims=['lat', 'lon'], coords=coords, iParams=dict(ip), mParams=mp)
        assert calls, "the device front end was not used"
        monkeypatch.setattr(apps, "_general_device_front", lambda *a, **k: None)
        s_h = fn(F, dims=['lat', 'lon'], coords=coords, iParams=dict(ip), mParams=mp)
        monkeypatch.undo()
        assert np.array_equal(s_f.values, s_h.values, equal_nan=True)
        lv = s_f.values[:, land]
        assert np.isnan(lv).all() if np.isnan(undef) else (lv == undef).all()
        assert np.isfinite(s_f.values[:, ~land]).all() and np.abs(s_f.values[:, ~land]).max() > 0


@pytest.mark.parametrize("n2kind", ["scalar", "profile_lev", "profile_lat", "volume", "full"])
@pytest.mark.parametrize("coords,bcs", [("lat-lon", ["fixed", "fixed", "periodic"]), ("lat-lon", ["fixed", "extend", "periodic"]),
                                        ("cartesian", ["fixed", "fixed", "fixed"])])
def test_omega_device_front_end_equals_host_path(gpu_ctx, monkeypatch, n2kind, coords, bcs):
    """invert_omega with icbc=None goes through xinv_std3d_rows (masking, A / B = N2*cosH / C = N2/cosG, forcing
    scale and de-masking on the device, N2 read through strides); the reference-shaped host path must give the same
    bits: every form of N2 the reference accepts, NaN-marked land, a batch over time."""
    from xinvert_b200 import apps
    nz, ny, nx, T = 9, 30, 64, 2
    lev = 100000.0 - 10000.0 * np.arange(nz)
    lat, lon = -58.0 + 4.0 * np.arange(ny), 5.625 * np.arange(nx)
    if coords == "cartesian":
        lat, lon = 1e5 * np.arange(ny) - 1.4e6, 1e5 * np.arange(nx)
    rng = np.random.default_rng(5)
    co = {'time': np.arange(T), 'LEV': lev, 'lat': lat, 'lon': lon}
    Fv = 1e-17 * rng.standard_normal((T, nz, ny, nx))
    Fv[:, 2:5, 8:12, 10:20] = np.nan                               # topography
    F = DA(Fv, ['time', 'LEV', 'lat', 'lon'], co)
    N2 = {"scalar": 2e-4,
          "profile_lev": DA(1e-6 * (1 + 0.5 * rng.random(nz)), ['LEV'], {'LEV': lev}),
          "profile_lat": DA(1e-6 * (1 + 0.5 * rng.random(ny)), ['lat'], {'lat': lat}),
          "volume": DA(1e-6 * (1 + 0.5 * rng.random((nz, ny, nx))), ['LEV', 'lat', 'lon'], {k: co[k] for k in ('LEV', 'lat', 'lon')}),
          "full": DA(1e-6 * (1 + 0.5 * rng.random((T, nz, ny, nx))), ['time', 'LEV', 'lat', 'lon'], co)}[n2kind]
    ip = {'BCs': bcs, 'tolerance': 1e-9, 'mxLoop': 40, 'printInfo': False}
    kw = dict(dims=['LEV', 'lat', 'lon'], coords=coords, mParams={'N2': N2, 'f0': 1e-4, 'beta': 2e-11})
    calls = []
    real = apps._device_solvers.solve_standard_3D_rows
    monkeypatch.setattr(apps._device_solvers, "solve_standard_3D_rows", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    ip_f = dict(ip)
    w_f = xb.invert_omega(F, iParams=ip_f, **kw)
    assert calls, "the device front end was not used"
    # N2 without a lon axis: A, B, C and the factor all as row values (X3_ROWS kernels)
    assert gpu_ctx.stats()["engine"] == "fused" and gpu_ctx.stats()["row_coeffs"] == (1 if n2kind in ("volume", "full") else 2)
    monkeypatch.setattr(apps, "_omega_device_front", lambda *a, **k: None)
    ip_h = dict(ip)
    w_h = xb.invert_omega(F, iParams=ip_h, **kw)
    assert np.array_equal(w_f.values, w_h.values, equal_nan=True)
    assert np.array_equal(ip_f['flags_all'], ip_h['flags_all'])
    assert np.isnan(w_f.values[:, 2:5, 8:12, 10:20]).all() and np.isfinite(w_f.values[:, 0]).all()
    assert np.abs(np.nan_to_num(w_f.values)).max() > 0


def test_omega_device_front_end_falls_back(gpu_ctx):
    """Odd nx with periodic-x is not a problem for the fused engine: the call still succeeds (host path, colour engine)."""
    nz, ny, nx = 6, 20, 31
    co = {'LEV': 100000.0 - 10000.0 * np.arange(nz), 'lat': -38.0 + 4.0 * np.arange(ny), 'lon': 360.0 / nx * np.arange(nx)}
    F = DA(1e-17 * np.random.default_rng(1).standard_normal((nz, ny, nx)), ['LEV', 'lat', 'lon'], co)
    ip = {'BCs': ['fixed', 'fixed', 'periodic'], 'tolerance': -1.0, 'mxLoop': 5, 'printInfo': False}
    w = xb.invert_omega(F, dims=['LEV', 'lat', 'lon'], iParams=ip, mParams={'N2': 2e-4})
    assert gpu_ctx.stats()["engine"] == "colour" and np.isfinite(w.values).all()
