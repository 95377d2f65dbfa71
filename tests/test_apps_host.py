"""CPU: host logic of the apps/core facade (masking, coefficient builders, grid
parameters, batching over non-core dims, de-masking, error behaviour), with the
C oracle standing in for the CUDA library (tests/oracle_backend.py is
monkeypatched into xinvert_b200.core; the product itself has no such path).

Known answers replayed here come from the reference's own tests
(tests/test_GillMatsuno.py:55-57, tests/test_Ishida.py:61-62,
tests/test_OptArg.py) and from reference-kernel probes recorded in SURVEY.md
section 4 / 8c; all use the reference's lexicographic ordering."""
import numpy as np
import pytest

import xinvert_b200 as xb
from xinvert_b200 import apps, core
from tests import oracle_backend

DA = xb.DataArray


@pytest.fixture(autouse=True)
def _oracle_backend(monkeypatch):
    monkeypatch.setattr(core, "solvers", oracle_backend)


def _grid2(ny, nx, y, x, ydim="lat", xdim="lon"):
    yy, xx = np.meshgrid(y, x, indexing="ij")
    return yy, xx, {ydim: y, xdim: x}


def test_gill_matsuno_idealized_known_answer(capsys):
    """tests/test_GillMatsuno.py:14-57 of the reference, verbatim numbers."""
    lon, lat = np.linspace(0, 360, 144), np.linspace(-90, 90, 73)
    la, lo, coords = _grid2(73, 144, lat, lon)
    Q1 = 0.05 * np.exp(-((la - 0) ** 2 + (lo - 120) ** 2) / 100.0)
    Q2 = 0.05 * np.exp(-((la - 10) ** 2 + (lo - 120) ** 2) / 100.0) - 0.05 * np.exp(-((la + 10) ** 2 + (lo - 120) ** 2) / 100.0)
    Q3 = 0.05 * np.exp(-((la - 10) ** 2 + (lo - 120) ** 2) / 100.0)
    iParams = {'BCs': ['fixed', 'periodic'], 'mxLoop': 2000, 'tolerance': 1e-8, 'optArg': 1.4,
               'ordering': 'lexicographic'}
    mParams = {'epsilon': 1e-5, 'Phi': 5000}
    want = [(4351.62244687, 1628), (5833.33192343, 1146), (5100.85325027, 1618)]
    hs = []
    for Q, (ke, loops) in zip((Q1, Q2, Q3), want):
        h = xb.invert_GillMatsuno(DA(Q, ['lat', 'lon'], coords), dims=['lat', 'lon'], iParams=iParams, mParams=mParams)
        assert h.name == 'inverted' and h.dims == ('lat', 'lon')
        u, v = xb.cal_flow(h, dims=['lat', 'lon'], BCs=['fixed', 'periodic'], mParams=mParams, vtype='GillMatsuno')
        assert np.isclose(((u ** 2 + v ** 2) / 2).sum(), ke, rtol=1e-10, atol=0)
        out = capsys.readouterr().out
        assert f"loops {loops:4.0f} and tolerance is" in out
        hs.append(h)
    assert (hs[0] <= 0).all() and (abs(hs[1]) <= 370).all() and (hs[2] <= 0).all()
    assert np.isclose(hs[0].min(), -364.71766191712015, rtol=1e-12)
    assert np.isclose(hs[2].min(), -539.8682258901698, rtol=1e-12)


def _ishida_case():
    xnum, ynum = 251, 151
    Lx, Ly = 1e7, 2 * np.pi * 1e6
    x, y = np.linspace(0, Lx, xnum), np.linspace(0, Ly, ynum)
    yg, xg, coords = _grid2(ynum, xnum, y, x, 'ydef', 'xdef')
    curl = -np.pi * np.sin(2. * np.pi * yg / Ly) / Ly
    return curl, coords, Ly


def test_ishida_land_mask_known_answer(capsys):
    """tests/test_Ishida.py:13-63 (general_2D with undef strips; user undef -9999)."""
    curl, coords, _ = _ishida_case()
    curl[65:, 100:104] = -9999
    curl[:75, 130:134] = -9999
    iParams = {'BCs': ['fixed', 'periodic'], 'mxLoop': 3000, 'tolerance': 1e-9, 'optArg': 1.4, 'undef': -9999,
               'ordering': 'lexicographic'}
    beta, R, depth = 2.2e-11, 0.0009, 200
    F = DA(curl, ['ydef', 'xdef'], coords)
    h1 = xb.invert_Stommel(F, dims=['ydef', 'xdef'], coords='cartesian', iParams=iParams,
                           mParams={'beta': beta, 'R': R, 'D': depth})
    out1 = capsys.readouterr().out
    h2 = xb.invert_Stommel(F, dims=['ydef', 'xdef'], coords='cartesian', iParams=iParams,
                           mParams={'beta': beta, 'R': R * 20, 'D': depth})
    out2 = capsys.readouterr().out
    land = (curl == -9999)
    assert (h1.values[land] == -9999).all() and (h2.values[land] == -9999).all()      # de-masked with the user's undef
    a1, a2 = np.abs(h1.values[~land]).max(), np.abs(h2.values[~land]).max()
    assert a1 <= 5.5e5 and a2 <= 2.8e4                                                # the reference's assertions
    assert "loops 1474" in out1 and np.isclose(a1, 451695.81539746444, rtol=1e-12)    # SURVEY.md section 4 probe
    assert "loops 3000" in out2 and np.isclose(a2, 26968.645791689938, rtol=1e-12)


def test_stommel_idealized_known_answer(capsys):
    """tests/test_StommelWBC.py:14-45, the general_2D halves (SURVEY.md section 4 probe values)."""
    xnum, ynum = 201, 151
    Lx, Ly = 1e7, 2 * np.pi * 1e6
    x, y = np.linspace(0, Lx, xnum), np.linspace(0, Ly, ynum)
    yg, xg, coords = _grid2(ynum, xnum, y, x, 'ydef', 'xdef')
    curl = DA(-0.3 * np.sin(np.pi * yg / Ly) * np.pi / Ly, ['ydef', 'xdef'], coords)
    iParams = {'BCs': ['fixed', 'fixed'], 'mxLoop': 5000, 'optArg': 1.9, 'tolerance': 1e-12,
               'ordering': 'lexicographic'}
    for beta, loops, mx in ((0, 3213, 611203.653077336), (1.8e-11, 457, 282080.3876195683)):
        S = xb.invert_Stommel(curl, dims=['ydef', 'xdef'], coords='cartesian', iParams=iParams,
                              mParams={'beta': beta, 'R': 0.0008, 'D': 200})
        assert f"loops {loops:4.0f}" in capsys.readouterr().out
        assert np.isclose(S.max(), mx, rtol=1e-12)


def _c1_zeta(ny=180, nx=360):
    lat = -90 + 90.0 / ny + (180.0 / ny) * np.arange(ny)
    lon = (360.0 / nx) * np.arange(nx)
    la, lo, coords = _grid2(ny, nx, lat, lon)
    lam, phi = np.deg2rad(lo), np.deg2rad(la)
    return 1e-5 * np.sin(3 * lam) * np.cos(phi) ** 2 * np.sin(2 * phi), coords


def test_poisson_c1_known_answer():
    """BASELINE configs[0] (SURVEY.md 8c KAT 6): the reference stops at loop 2380 with
    rel-change 9.99090342e-09 and max|psi| = 13182413.993245527."""
    zeta, coords = _c1_zeta()
    ip = {'BCs': ['fixed', 'periodic'], 'optArg': 1.4, 'tolerance': 1e-8, 'mxLoop': 5000, 'printInfo': False,
          'ordering': 'lexicographic'}
    psi = xb.invert_Poisson(DA(zeta, ['lat', 'lon'], coords), dims=['lat', 'lon'], iParams=ip)
    assert np.isclose(np.abs(psi.values).max(), 13182413.993245527, rtol=1e-12)
    assert 'flags' not in ip                       # the caller's dict is never mutated (apps.py:1361 deep-copies)


def test_poisson_batched_time_axis_demask_and_icbc(capsys):
    """Non-core dims are flattened into one batched call; each slice equals its own
    2-D solve; undef cells come back as iParams['undef']; icbc keeps edge/land values."""
    zeta, coords = _c1_zeta(30, 48)
    T = 3
    z3 = np.stack([zeta * (1 + t) for t in range(T)])
    z3[:, 10:14, 20:25] = np.nan                                   # land (default undef = nan)
    c3 = dict(coords, time=np.arange(T))
    ip = {'BCs': ['extend', 'periodic'], 'tolerance': 1e-7, 'mxLoop': 400}
    out = xb.invert_Poisson(DA(z3, ['time', 'lat', 'lon'], c3), dims=['lat', 'lon'], iParams=ip)
    lines = capsys.readouterr().out.strip().splitlines()
    assert len(lines) == T and lines[0].startswith("{time: 0}") and " loops " in lines[0]
    assert out.shape == (T, 30, 48) and np.isnan(out.values[:, 10:14, 20:25]).all()
    for t in range(T):
        one = xb.invert_Poisson(DA(z3[t], ['lat', 'lon'], coords), dims=['lat', 'lon'], iParams=dict(ip, printInfo=False))
        assert np.array_equal(out.values[t], one.values, equal_nan=True)
    # time axis in the middle of the dims is handled too (core dims keep their order)
    mid = xb.invert_Poisson(DA(np.moveaxis(z3, 0, 1), ['lat', 'time', 'lon'], c3), dims=['lat', 'lon'],
                            iParams=dict(ip, printInfo=False))
    assert np.array_equal(np.moveaxis(mid.values, 1, 0), out.values, equal_nan=True)
    # icbc: fixed-y edge rows and land cells take the prescribed values, interior starts from 0
    icbc = DA(np.full((30, 48), 7.0), ['lat', 'lon'], coords)
    r = xb.invert_Poisson(DA(z3[0], ['lat', 'lon'], coords), dims=['lat', 'lon'], icbc=icbc,
                          iParams={'BCs': ['fixed', 'periodic'], 'mxLoop': 5, 'printInfo': False})
    assert (r.values[0] == 7.0).all() and (r.values[-1] == 7.0).all() and (r.values[10:14, 20:25] == 7.0).all()
    assert not (r.values[5] == 7.0).any()


def test_omega_coefficients_and_solve():
    """invert_omega: lat-lon coefficients as in apps.py:2025-2036, checked against a
    direct oracle call with hand-built arrays; N2 given as a level profile."""
    nz, ny, nx = 7, 16, 24
    lev = 100000.0 - 12500.0 * np.arange(nz)
    lat = -60 + 8.0 * np.arange(ny)
    lon = 15.0 * np.arange(nx)
    rng = np.random.default_rng(3)
    F = 1e-17 * rng.standard_normal((nz, ny, nx))
    N2 = 1e-6 * (1 + 0.5 * rng.random(nz))
    coords = {'lev': lev, 'lat': lat, 'lon': lon}
    ip = {'BCs': ['fixed', 'fixed', 'periodic'], 'tolerance': 1e-9, 'mxLoop': 300, 'printInfo': False,
          'ordering': 'lexicographic'}
    w = xb.invert_omega(DA(F, ['lev', 'lat', 'lon'], coords), dims=['lev', 'lat', 'lon'], iParams=ip,
                        mParams={'N2': DA(N2, ['lev'], {'lev': lev})})
    import oracle
    lats = np.deg2rad(lat)
    cosG = np.cos(lats)
    cosH = np.cos((lats + np.r_[np.nan, lats[:-1]]) / 2)
    f = 2 * 7.292e-5 * np.sin(lats)
    A = np.broadcast_to((f ** 2 * cosG)[None, :, None], F.shape).copy()
    B = (N2[:, None, None] * cosH[None, :, None] * np.ones(F.shape)).copy()
    C = (N2[:, None, None] / cosG[None, :, None] * np.ones(F.shape)).copy()
    Fm = (F * cosG[None, :, None]).copy()
    Re = 6371200.0
    d3, d2, d1 = -12500.0, np.deg2rad(8.0) * Re, np.deg2rad(15.0) * Re
    eps = np.sin(np.pi / (2. * nx + 2)) ** 2 + np.sin(np.pi / (2. * ny + 2)) ** 2 + np.sin(np.pi / (2. * nz + 3)) ** 2
    S = np.zeros(F.shape)
    fl = np.array([0., 1., 0.])
    oracle.invert_standard_3D(S, A, B, C, Fm, nz, ny, nx, d3, d2, d1, 'fixed', 'fixed', 'periodic', d1 ** 2,
                              (d1 / d3) ** 2, (d1 / d2) ** 2, 2 / (1 + np.sqrt((2 - eps) * eps)), -9.99e8, fl, 300, 1e-9)
    assert np.array_equal(w.values, S)
    with pytest.raises(Exception, match="unstable stratification"):
        xb.invert_omega(DA(F, ['lev', 'lat', 'lon'], coords), dims=['lev', 'lat', 'lon'], iParams=ip,
                        mParams={'N2': DA(-N2, ['lev'], {'lev': lev})})


def test_eliassen_nine_point_uses_B():
    ny, nx = 20, 30
    z, y = np.linspace(1000., 100., ny), np.linspace(0., 5e5, nx)
    coords = {'z': z, 'r': y}
    rng = np.random.default_rng(5)
    A = DA(1 + 0.2 * rng.random((ny, nx)), ['z', 'r'], coords)
    B = DA(0.1 * rng.standard_normal((ny, nx)), ['z', 'r'], coords)
    C = DA(1 + 0.2 * rng.random((ny, nx)), ['z', 'r'], coords)
    F = DA(1e-9 * rng.standard_normal((ny, nx)), ['z', 'r'], coords)
    ip = {'BCs': ['fixed', 'fixed'], 'mxLoop': 50, 'tolerance': 1e-12, 'optArg': 1.2, 'printInfo': False}
    s1 = xb.invert_Eliassen(F, dims=['z', 'r'], coords='cartesian', iParams=ip, mParams={'A': A, 'B': B, 'C': C})
    s0 = xb.invert_Eliassen(F, dims=['z', 'r'], coords='cartesian', iParams=ip, mParams={'A': A, 'B': B * 0, 'C': C})
    assert s1.shape == (ny, nx) and np.isfinite(s1.values).all() and not np.allclose(s1.values, s0.values)
    with pytest.raises(Exception, match="unsupported coords"):
        xb.invert_Eliassen(F, dims=['z', 'r'], coords='lat-lon', iParams=ip, mParams={'A': A, 'B': B, 'C': C})


def test_parameter_handling_and_errors():
    zeta, coords = _c1_zeta(12, 16)
    F = DA(zeta, ['lat', 'lon'], coords)
    with pytest.raises(Exception, match="dimensional forcing"):
        xb.invert_Poisson(F, dims=['lat'])
    with pytest.raises(Exception, match="is not used"):
        xb.invert_Poisson(F, dims=['lat', 'lon'], mParams={'Phi': 1.0})
    bad = dict(coords, lon=np.r_[coords['lon'][:-1], 400.0])
    with pytest.raises(Exception, match="non-uniform"):
        xb.invert_Poisson(DA(zeta, ['lat', 'lon'], bad), dims=['lat', 'lon'])
    with pytest.raises(Exception, match="unsupported coords"):
        xb.invert_Poisson(F, dims=['lat', 'lon'], coords='polar')
    # tests/test_OptArg.py: 1 <= optimal omega <= 2 for any grid
    for n in (3, 10, 100, 4000):
        g = apps._Grid(DA(np.zeros((n, n + 1)), ['y', 'x'], {'y': np.arange(n) * 1.0, 'x': np.arange(n + 1) * 2.0}), ['y', 'x'])
        p = apps._cal_params2D(g, 'cartesian', 6371200.0)
        assert 1.0 <= p['optArg'] <= 2.0 and p['ratio'] == 2.0 and p['ratioQtr'] == 0.5 and p['del1Sqr'] == 4.0
    # None in iParams does not override the computed default (apps.py:2371-2373)
    r = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams={'optArg': None, 'mxLoop': 3, 'printInfo': False,
                                                           'BCs': ['fixed', 'periodic']})
    assert r.shape == (12, 16)


def test_non_finite_forcing_builds_full_size_coefficients():
    """inf in F makes the reference's zero = maskF - maskF NaN there; the facade then
    builds per-slice coefficients like the reference instead of shared ones."""
    zeta, coords = _c1_zeta(12, 16)
    z = np.stack([zeta, zeta])
    z[1, 5, 5] = np.inf
    g = apps._Grid(DA(z, ['t', 'lat', 'lon'], dict(coords, t=np.arange(2))), ['lat', 'lon'])
    ip = apps._update(apps.default_iParams, {})
    maskF, Fm, S0, (A, B, C) = apps._coeffs_Poisson(g, 'lat-lon', apps.default_mParams, ip, None)
    assert A.values.shape == z.shape and np.isnan(A.values[1, 5, 5]) and np.isfinite(A.values[0, 1:]).all()
    z[1, 5, 5] = 1.0
    g = apps._Grid(DA(z, ['t', 'lat', 'lon'], dict(coords, t=np.arange(2))), ['lat', 'lon'])
    _, _, _, (A, B, C) = apps._coeffs_Poisson(g, 'lat-lon', apps.default_mParams, ip, None)
    assert A.values.shape == (12, 16) and not B.values.any()
