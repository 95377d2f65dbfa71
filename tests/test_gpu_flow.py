"""GPU: cal_flow's array work on the device (xinv_flow2d) equals the numpy evaluation of the reference's expressions
(apps._flow_host) BIT FOR BIT -- both branches, every boundary padding, uniform and non-uniform coordinates,
NaN-marked land, batches -- and reproduces the reference's Gill-Matsuno kinetic-energy goldens."""
import numpy as np
import pytest

import xinvert_b200 as xb
from xinvert_b200 import apps, solvers

pytestmark = pytest.mark.gpu
DA = xb.DataArray


def _S(ny, nx, T, coords, kind):
    rng = np.random.default_rng(ny * 1000 + nx)
    if coords == 'lat-lon':
        lat = {"uniform": -59.0 + 2.0 * np.arange(ny), "linspace": np.linspace(-80, 80, ny)}[kind]
        lon = {"uniform": 2.5 * np.arange(nx), "linspace": np.linspace(0, 357.5, nx)}[kind]
    else:
        lat = {"uniform": 1e5 * np.arange(ny) - 2e6, "linspace": np.cumsum(1e5 * (1 + 0.2 * rng.random(ny)))}[kind]
        lon = {"uniform": 1e5 * np.arange(nx), "linspace": np.cumsum(1e5 * (1 + 0.2 * rng.random(nx)))}[kind]
    v = rng.standard_normal((T, ny, nx)) * 1e5
    v[:, 5:9, 10:17] = np.nan                                     # land
    return DA(v, ['time', 'lat', 'lon'], {'time': np.arange(T), 'lat': lat, 'lon': lon})


def _host(S, dims, **kw):
    """The same call with the numpy path forced."""
    from xinvert_b200 import core
    from tests import oracle_backend
    saved, core.solvers = core.solvers, oracle_backend
    try:
        return xb.cal_flow(S, dims, **kw)
    finally:
        core.solvers = saved


@pytest.mark.parametrize("kind", ["uniform", "linspace"])
@pytest.mark.parametrize("coords", ["lat-lon", "cartesian"])
@pytest.mark.parametrize("vtype", ["streamfunction", "velocitypotential"])
@pytest.mark.parametrize("BCs", [["fixed", "fixed"], ["extend", "periodic"], ["reflect", "extend"], ["fixed", "periodic"]])
def test_flow_from_streamfunction_device_equals_numpy(gpu_ctx, monkeypatch, kind, coords, vtype, BCs):
    S = _S(33, 70, 3, coords, kind)
    calls = []
    real = solvers.flow_2d
    monkeypatch.setattr(apps._device_solvers, "flow_2d", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    u, v = xb.cal_flow(S, ['lat', 'lon'], coords=coords, BCs=BCs, vtype=vtype)
    assert calls, "the device path was not used"
    uh, vh = _host(S, ['lat', 'lon'], coords=coords, BCs=BCs, vtype=vtype)
    assert np.array_equal(u.values, uh.values, equal_nan=True) and np.array_equal(v.values, vh.values, equal_nan=True)
    assert np.isfinite(u.values[:, 20:, 30:]).all() and np.abs(u.values[:, 20:, 30:]).max() > 0


@pytest.mark.parametrize("kind", ["uniform", "linspace"])
@pytest.mark.parametrize("coords", ["lat-lon", "cartesian"])
def test_flow_gill_matsuno_device_equals_numpy(gpu_ctx, kind, coords):
    S = _S(40, 64, 2, coords, kind)
    mp = {'f0': 1e-5, 'beta': 2e-11, 'epsilon': 1e-5, 'Phi': 5000}
    u, v = xb.cal_flow(S, ['lat', 'lon'], coords=coords, vtype='GillMatsuno', mParams=mp)
    uh, vh = _host(S, ['lat', 'lon'], coords=coords, vtype='GillMatsuno', mParams=mp)
    assert np.array_equal(u.values, uh.values, equal_nan=True) and np.array_equal(v.values, vh.values, equal_nan=True)


def test_flow_falls_back_to_numpy_for_non_trailing_dims_and_float32(gpu_ctx):
    S = _S(20, 36, 2, 'lat-lon', 'uniform')
    St = DA(np.ascontiguousarray(np.moveaxis(S.values, 0, 2)), ['lat', 'lon', 'time'], S.coords)
    u, v = xb.cal_flow(St, ['lat', 'lon'], BCs=['extend', 'periodic'])
    u0, v0 = xb.cal_flow(S, ['lat', 'lon'], BCs=['extend', 'periodic'])
    assert np.array_equal(np.moveaxis(u.values, 2, 0), u0.values, equal_nan=True)
    S32 = DA(S.values.astype(np.float32), S.dims, S.coords)
    u32, _ = xb.cal_flow(S32, ['lat', 'lon'], BCs=['extend', 'periodic'])
    assert np.isfinite(u32.values[:, 12:, 20:]).all()
