"""cal_flow on the host (numpy restatement of apps.py:1181-1317 and of the finite-difference pieces it uses,
finitediffs.py:151-207, :548-659): CPU-only checks of the branches the GPU path mirrors.

* ``axis_diff`` describes numpy.gradient's formulas: an explicit per-index evaluation of the description must give
  the bits numpy gives (uniform and non-uniform coordinates, every padding mode);
* the physical identity the reference's own test uses (tests/test_Poisson.py:37-41): the rotational flow of a
  streamfunction is non-divergent, the divergent flow of a potential is irrotational (discretely, with periodic-x);
* 'streamfunction' and 'velocitypotential' return the same two gradients, ordered and signed as apps.py:1221-1224.
"""
import numpy as np
import pytest

import xinvert_b200 as xb
from xinvert_b200 import apps, core, solvers
from tests import oracle_backend

DA = xb.DataArray


@pytest.fixture(autouse=True)
def _host_backend(monkeypatch):
    monkeypatch.setattr(core, "solvers", oracle_backend)       # cal_flow then takes its numpy path


def _explicit(f, d):
    """axis_diff's description evaluated index by index along the last axis (what the CUDA kernel does)."""
    n = f.shape[-1]
    out = np.empty_like(f)
    for i in range(n):
        if d["edge"] is None and i == 0:
            out[..., i] = (f[..., 1] - f[..., 0]) / d["lo"]
            continue
        if d["edge"] is None and i == n - 1:
            out[..., i] = (f[..., n - 1] - f[..., n - 2]) / d["hi"]
            continue
        e = d["edge"]
        fm = f[..., i - 1] if i > 0 else {"fixed": np.full(f.shape[:-1], d["lo"]), "extend": f[..., 0], "reflect": f[..., 1],
                                          "periodic": f[..., n - 1]}[e]
        fp = f[..., i + 1] if i < n - 1 else {"fixed": np.full(f.shape[:-1], d["hi"]), "extend": f[..., n - 1],
                                              "reflect": f[..., n - 2], "periodic": f[..., 0]}[e]
        if d["uniform"]:
            out[..., i] = (fp - fm) / d["den"]
        else:
            out[..., i] = (d["w"][0, i] * fm + d["w"][1, i] * f[..., i]) + d["w"][2, i] * fp
    return out


@pytest.mark.parametrize("edge", [None, "fixed", "extend", "reflect", "periodic"])
@pytest.mark.parametrize("coord", ["uniform_exact", "linspace", "stretched"])
def test_axis_diff_describes_numpy_gradient(edge, coord):
    n = 37
    x = {"uniform_exact": 2.5 * np.arange(n), "linspace": np.linspace(-90, 90, n),
         "stretched": np.cumsum(1.0 + 0.3 * np.random.default_rng(1).random(n))}[coord]
    f = np.random.default_rng(2).standard_normal((4, n))
    d = solvers.axis_diff(x, edge, fill=(0.25, -1.5))
    assert d["uniform"] == (coord == "uniform_exact" or (coord == "linspace" and bool((np.diff(
        x if edge is None else np.concatenate([[2 * x[0] - x[1]], x, [2 * x[-1] - x[-2]]])) == np.diff(x)[0]).all())))
    want = apps._np_diff(f, x, 1, d)
    assert np.array_equal(_explicit(f, d), want)


def _field(ny=40, nx=72, T=2):
    lat, lon = -58.5 + 3.0 * np.arange(ny), 5.0 * np.arange(nx)
    lam, phi = np.deg2rad(lon)[None, :], np.deg2rad(lat)[:, None]
    base = 1e6 * np.sin(2 * lam) * np.cos(phi) ** 2 + 3e5 * np.cos(3 * lam + 0.3) * np.sin(2 * phi)
    v = np.stack([base * (1 + 0.1 * t) for t in range(T)])
    return DA(v, ['time', 'lat', 'lon'], {'time': np.arange(T), 'lat': lat, 'lon': lon}), lat, lon


def test_streamfunction_and_velocitypotential_return_the_same_gradients():
    S, lat, lon = _field()
    u, v = xb.cal_flow(S, ['lat', 'lon'], BCs=['extend', 'periodic'], vtype='streamfunction')
    a, b = xb.cal_flow(S, ['lat', 'lon'], BCs=['extend', 'periodic'], vtype='velocitypotential')
    assert np.array_equal(u.values, -b.values) and np.array_equal(v.values, a.values)     # (-grdy, grdx) vs (grdx, grdy)
    # d/dy of the streamfunction against a plain centred difference, metric pi R / 180
    deg2m = np.pi * 6371200.0 / 180.0
    want = -(S.values[:, 2:, :] - S.values[:, :-2, :]) / (lat[2:] - lat[:-2])[None, :, None] / deg2m
    assert np.allclose(u.values[:, 1:-1, :], want, rtol=1e-12)


def test_rotational_flow_is_nondivergent_on_a_cartesian_periodic_grid():
    """u = -dpsi/dy, v = dpsi/dx with centred differences: du/dx + dv/dy vanishes identically inside (the mixed
    differences commute) -- the identity tests/test_Poisson.py:37-41 checks on real data."""
    ny, nx = 30, 48
    y, x = 1e5 * np.arange(ny), 1e5 * np.arange(nx)
    rng = np.random.default_rng(3)
    psi = DA(rng.standard_normal((ny, nx)), ['y', 'x'], {'y': y, 'x': x})
    u, v = xb.cal_flow(psi, ['y', 'x'], coords='cartesian', BCs=['extend', 'periodic'], vtype='streamfunction')
    dudx = (np.roll(u.values, -1, 1) - np.roll(u.values, 1, 1)) / 2e5
    dvdy = (v.values[2:] - v.values[:-2]) / 2e5
    div = dudx[1:-1] + dvdy
    assert np.abs(div[1:-1]).max() < 1e-12 * np.abs(u.values).max() / 1e5 * 1e5


def test_z_plane_variants_and_errors():
    nz, ny = 8, 20
    S = DA(np.random.default_rng(4).standard_normal((nz, ny)), ['lev', 'lat'], {'lev': 1000.0 * np.arange(nz), 'lat': -47.5 + 5.0 * np.arange(ny)})
    a, b = xb.cal_flow(S, ['lev', 'lat'], coords='z-lat', BCs=['extend', 'extend'])
    assert a.values.shape == (nz, ny) and np.isfinite(a.values).all() and np.isfinite(b.values).all()
    with pytest.raises(Exception, match='unsupported vtype'):
        xb.cal_flow(S, ['lev', 'lat'], vtype='vorticity')
    with pytest.raises(Exception, match='unsupported coords'):
        xb.cal_flow(S, ['lev', 'lat'], coords='polar')
