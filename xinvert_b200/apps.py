"""Drop-in for the five hot-path entry points of ``xinvert/apps.py``:
``invert_Poisson``, ``invert_Eliassen``, ``invert_GillMatsuno``,
``invert_Stommel`` and ``invert_omega`` (apps.py:67-100, :300-346, :349-394,
:445-488, :766-827), same arguments, defaults, errors and return value
(a DataArray named ``'inverted'``), plus ``cal_flow(..., vtype='GillMatsuno')``
(apps.py:1277-1317).

The bodies restate the reference's ``__template`` / ``__mask_FS`` /
``__cal_params2D/3D`` / ``__coeffs_*`` / ``__update`` (apps.py:1324-1394,
:2112-2159, :2162-2313, :1397-1437, :1582-1657, :1712-1748, :2016-2052,
:2361-2379) in plain numpy -- ``apps.py`` itself cannot be imported without
xarray -- and hand the arrays to ``xinvert_b200.core`` (one batched C-ABI call
into the CUDA library instead of a serial loop over numba kernels).

Differences from the reference, all deliberate:
* coefficients that do not depend on the non-core dims (e.g. ``cosH``,
  ``1/cosG``, ``c1*Phi``) are built ONCE with the core shape and shared by the
  whole batch (stride 0) instead of being materialised per slice; the values
  each slice sees are identical.  (If the forcing holds non-finite values that
  the reference would smear into its ``zero = maskF - maskF`` template, the
  full-size arrays are built instead, so that even those cases agree.)
* the default update order is red-black ('colour'); ``iParams['ordering'] =
  'lexicographic'`` selects the reference's own order (see core.py's docstring).
* inputs are promoted to float64 for the solve (the reference iterates float32
  inputs with float32 stores); the result is cast back to the input dtype.
"""
import copy

import numpy as np

from . import core
from . import solvers as _device_solvers
from ._lib import E_UNSUPPORTED, XinvError
from .xrshim import coord_values, wrap_like

_undeftmp = -9.99e8

###### default invert parameters (apps.py:21-38) ######
default_iParams = copy.deepcopy({
    'BCs'      : ['fixed', 'fixed'],
    'undef'    : np.nan,
    'mxLoop'   : 5000,
    'tolerance': 1e-8,
    'optArg'   : None,
    'printInfo': True,
    'debug'    : False,
})

###### default model parameters (apps.py:42-60) ######
default_mParams = copy.deepcopy({
    'f0'     : 1e-5 ,
    'beta'   : 2e-11,
    'Phi'    : 1e4  ,
    'epsilon': 7e-6 ,
    'N2'     : 2e-4 ,
    'A'      : 1e5  ,
    'R'      : 5e-5 ,
    'depth'  : 100  ,
    'rho0'   : 1027 ,
    'ang0'   : 2e5  ,
    'lambda' : 1e-8 ,
    'c0'     : 8e-9 ,
    'c1'     : 8e-5 ,
    'Rearth' : 6371200.0,
    'Omega'  : 7.292e-5 ,
    'g'      : 9.80665  ,
})


# ---------------------------------------------------------------------------
# application functions
# ---------------------------------------------------------------------------
def invert_Poisson(F, dims, coords='lat-lon', icbc=None,
                   mParams=default_mParams, iParams=default_iParams):
    r"""Invert :math:`\psi_{yy} + \psi_{xx} = F` for :math:`\psi` (apps.py:67-100).

    With ``icbc=None`` (the common case) the masking, the coefficient arrays and the de-masking are
    done on the device (``xinv_std2d_rows``) from the forcing and three per-row vectors -- same
    numbers, no full-size host temporaries; anything else takes the reference-shaped host path."""
    fast = _poisson_device_front(F, dims, coords, icbc, mParams, iParams)
    if fast is not None:
        return fast
    return _template(_coeffs_Poisson, core.inv_standard2D, 2, F, dims, coords,
                     icbc, ['g', 'Omega', 'Rearth'], mParams, iParams)


def invert_Eliassen(F, dims, coords='z-lat', icbc=None,
                    mParams=default_mParams, iParams=default_iParams):
    r"""Invert the Eliassen balanced-vortex equation with user-supplied
    ``mParams['A'|'B'|'C']`` fields (apps.py:300-346); 9-point stencil.  With ``icbc=None`` the masks, the zero
    initial guess and the de-masking are done on the device (``xinv_std2d_front``)."""
    fast = _eliassen_device_front(F, dims, coords, icbc, mParams, iParams)
    if fast is not None:
        return fast
    return _template(_coeffs_Eliassen, core.inv_standard2D, 2, F, dims, coords,
                     icbc, ['A', 'B', 'C', 'g', 'Omega', 'Rearth'], mParams, iParams)


def invert_GillMatsuno(Q, dims, coords='lat-lon', icbc=None,
                       mParams=default_mParams, iParams=default_iParams):
    r"""Invert the Gill-Matsuno model for the mass field :math:`\phi` given the
    heating ``Q`` (apps.py:349-394); winds follow from ``cal_flow``."""
    valid = ['f0', 'beta', 'epsilon', 'Phi', 'g', 'Omega', 'Rearth']
    fast = _general_device_front(_rows_GillMatsuno, valid, Q, dims, coords, icbc, mParams, iParams)
    if fast is not None:
        return fast
    return _template(_coeffs_GillMatsuno, core.inv_general2D, 2, Q, dims, coords,
                     icbc, ['f0', 'beta', 'epsilon', 'Phi', 'g', 'Omega', 'Rearth'],
                     mParams, iParams)


def invert_Stommel(curl, dims, coords='lat-lon', icbc=None,
                   mParams=default_mParams, iParams=default_iParams):
    r"""Invert the Stommel model for the streamfunction given the wind-stress
    curl (apps.py:445-488)."""
    valid = ['beta', 'R', 'D', 'rho0', 'g', 'Omega', 'Rearth']
    fast = _general_device_front(_rows_Stommel, valid, curl, dims, coords, icbc, mParams, iParams)
    if fast is not None:
        return fast
    return _template(_coeffs_Stommel, core.inv_general2D, 2, curl, dims, coords,
                     icbc, ['beta', 'R', 'D', 'rho0', 'g', 'Omega', 'Rearth'],
                     mParams, iParams)


def invert_omega(F, dims, coords='lat-lon', icbc=None,
                 mParams=default_mParams, iParams=default_iParams):
    r"""Invert the quasi-geostrophic omega equation in 3-D (apps.py:766-827)."""
    N2 = mParams['N2'] if 'N2' in mParams else None
    if hasattr(N2, 'dims'):                      # array-valued stratification (apps.py:815-823)
        v = np.asarray(N2.values)[1:]
        if not np.isfinite(v).all():
            raise Exception('inifinite stratification coefficient A')
        if np.isnan(v).any():
            raise Exception('nan in coefficient A')
        if (v <= 0).any():
            raise Exception('unstable stratification in coefficient A')
    fast = _omega_device_front(F, dims, coords, icbc, mParams, iParams)
    if fast is not None:
        return fast
    return _template(_coeffs_omega, core.inv_standard3D, 3, F, dims, coords,
                     icbc, ['f0', 'beta', 'N2', 'g', 'Omega', 'Rearth'],
                     mParams, iParams)


def cal_flow(S, dims, coords='lat-lon', BCs=['fixed', 'fixed'],
             vtype='streamfunction', mParams=default_mParams):
    """Flow vector from the inverted field (apps.py:1181-1317): (u, v) from a streamfunction or a velocity
    potential (centred differences with the boundary conditions padded on, divided by the metric;
    apps.py:1207-1271, finitediffs.py:151-207, :548-659), or from the Gill-Matsuno mass field
    (apps.py:1277-1317).

    The array work runs on the device (``xinv_flow2d``: the same IEEE operations, in the order of the numpy
    expressions the reference evaluates) when ``S`` is float64 with the two core dims last; otherwise -- and for
    the 'z-lat' / 'z-lon' variants -- in numpy on the host (``_flow_host``), which gives the same bits."""
    if vtype.lower() not in ['streamfunction', 'velocitypotential', 'gillmatsuno']:
        raise Exception('unsupported vtype: ' + vtype + ', should be one of:\n' +
                        "['streamfunction', 'velocitypotential', 'gillmatsuno']")
    if len(dims) != 2:
        raise Exception('2 dimensions are needed')
    sv = np.asarray(S.values)
    ay, ax = list(S.dims).index(dims[0]), list(S.dims).index(dims[1])
    ydef = np.asarray(coord_values(S, dims[0]), dtype=np.float64)
    xdef = np.asarray(coord_values(S, dims[1]), dtype=np.float64)
    ny = ydef.size
    c = coords.lower()
    if vtype != 'GillMatsuno':                    # Poisson case
        sf = (vtype == 'streamfunction')
        R = 6371200.0                             # FiniteDiff's default radius (finitediffs.py:40)
        deg2m = np.pi * R / 180.0
        if c == 'lat-lon':
            rows = np.stack([np.full(ny, deg2m), deg2m * np.cos(np.deg2rad(ydef))])
        elif c == 'cartesian':
            rows = np.ones((2, ny))
        elif c in ('z-lat', 'z-lon'):
            return _flow_zplane(S, sv, ay, ax, ydef, xdef, c, BCs, sf, deg2m)
        else:
            raise Exception('unsupported coords ' + coords + ', should be [lat-lon, z-lat, z-lon, cartesian]')
        plan = dict(ydiff=_device_solvers.axis_diff(ydef, BCs[0]), xdiff=_device_solvers.axis_diff(xdef, BCs[1]),
                    comb=0, rows=rows, swap=not sf, signs=(-1.0, 1.0) if sf else (1.0, 1.0))
        for b in BCs:
            if b not in ('fixed', 'extend', 'reflect', 'periodic'):
                raise Exception('unsupported BC: ' + str(BCs))
    else:                                         # GillMatsuno case
        mParams = _update(default_mParams, mParams, ['f0', 'beta', 'epsilon', 'Phi', 'Omega', 'Rearth'])
        eps, f0, beta = mParams['epsilon'], mParams['f0'], mParams['beta']
        Omega, Rearth = mParams['Omega'], mParams['Rearth']
        # DataArray.differentiate == np.gradient along the coordinate (2nd order inside, 1st at the edges)
        yd, xd = _device_solvers.axis_diff(ydef, None), _device_solvers.axis_diff(xdef, None)
        if c == 'lat-lon':
            lats = np.deg2rad(ydef)
            f = 2.0 * Omega * np.sin(lats)
            rows = np.stack([eps / (eps ** 2.0 + f ** 2.0), f / (eps ** 2.0 + f ** 2.0), np.cos(lats)])
            plan = dict(ydiff=yd, xdiff=xd, comb=1, rows=rows, deg2m=np.deg2rad(1.0) * Rearth)
        elif c == 'cartesian':
            f = f0 + beta * ydef
            rows = np.stack([eps / (eps ** 2.0 + f ** 2.0), f / (eps ** 2.0 + f ** 2.0)])
            plan = dict(ydiff=yd, xdiff=xd, comb=2, rows=rows)
        else:
            raise Exception('unsupported coords ' + coords + ', should be [lat-lon, cartesian]')
    trailing = (ay == sv.ndim - 2 and ax == sv.ndim - 1)
    if trailing and sv.dtype == np.float64 and core.solvers is _device_solvers:
        c1, c2 = _device_solvers.flow_2d(sv, **plan)
    else:
        c1, c2 = _flow_host(sv, ay, ax, ydef, xdef, **plan)
    return wrap_like(S, c1), wrap_like(S, c2)


def _np_diff(sv, coord, axis, d):
    """numpy's own evaluation of what ``axis_diff`` describes: np.gradient of the (padded) array along ``axis``."""
    if d["edge"] is None:
        return np.gradient(sv, coord, axis=axis, edge_order=1)
    pw = [(0, 0)] * sv.ndim
    pw[axis] = (1, 1)
    if d["edge"] == 'periodic':
        p = np.pad(sv, pw, mode='wrap')
    elif d["edge"] == 'fixed':
        p = np.pad(sv, pw, mode='constant', constant_values=[(0, 0)] * axis + [(d["lo"], d["hi"])] + [(0, 0)] * (sv.ndim - axis - 1))
    else:
        p = np.pad(sv, pw, mode={'extend': 'edge', 'reflect': 'reflect'}[d["edge"]])
    x = np.concatenate([[0.0], np.asarray(coord, dtype=np.float64), [0.0]])
    x[0] = x[1] * 2 - x[2]
    x[-1] = x[-2] * 2 - x[-3]
    g = np.gradient(p, x, axis=axis, edge_order=1)
    sl = [slice(None)] * sv.ndim
    sl[axis] = slice(1, -1)
    return g[tuple(sl)]


def _flow_host(sv, ay, ax, ydef, xdef, ydiff, xdiff, comb, rows, swap=False, signs=(1.0, 1.0), deg2m=1.0):
    """The host (numpy) evaluation of cal_flow's array work: the expressions of apps.py:1207-1317 as they stand."""
    sv = np.asarray(sv, dtype=np.float64)
    dSdy = _np_diff(sv, ydef, ay, ydiff)
    dSdx = _np_diff(sv, xdef, ax, xdiff)

    def along_y(v):                               # broadcast a function of dims[0] against S
        shp = [1] * sv.ndim
        shp[ay] = -1
        return np.reshape(v, shp)

    if comb == 0:
        gy, gx = dSdy / along_y(rows[0]), dSdx / along_y(rows[1])
        r1, r2 = (gx, gy) if swap else (gy, gx)
        return (-r1 if signs[0] < 0 else r1), (-r2 if signs[1] < 0 else r2)
    coef1, coef2 = along_y(rows[0]), along_y(rows[1])
    if comb == 1:
        cosLat = along_y(rows[2])
        c1 = - coef1 * dSdx / deg2m / cosLat - coef2 * dSdy / deg2m
        c2 = - coef1 * dSdy / deg2m + coef2 * dSdx / deg2m / cosLat
    else:
        c1 = - coef1 * dSdx - coef2 * dSdy
        c2 = - coef1 * dSdy + coef2 * dSdx
    return c1, c2


def _flow_zplane(S, sv, ay, ax, ydef, xdef, c, BCs, sf, deg2m):
    """The 'z-lat' / 'z-lon' variants of the streamfunction / velocity-potential branch (apps.py:1225-1256), numpy."""
    sv = np.asarray(sv, dtype=np.float64)
    d0, d1 = _device_solvers.axis_diff(ydef, BCs[0]), _device_solvers.axis_diff(xdef, BCs[1])
    g0 = _np_diff(sv, ydef, ay, d0)               # d/dz: scale 1
    g1 = _np_diff(sv, xdef, ax, d1) / deg2m       # d/dlat or d/dlon with no latitude dimension: scale deg2m (cos = 1)
    if c == 'z-lat':
        shp = [1] * sv.ndim
        shp[ax] = -1
        cos = np.reshape(np.cos(np.deg2rad(xdef)), shp)
        g0, g1 = g0 / cos, g1 / cos
        g1 = np.where(np.reshape(np.abs(xdef), shp) != 90, g1, 0)
        r = (-g0, g1) if sf else (g1, g0)
    else:
        r = (g0, -g1) if sf else (g1, g0)
    return wrap_like(S, r[0]), wrap_like(S, r[1])


# ---------------------------------------------------------------------------
# helpers (numpy restatement of the reference's private functions)
# ---------------------------------------------------------------------------
class _Grid:
    """The forcing as numpy + the axis bookkeeping every coefficient builder needs."""

    def __init__(self, F, dims):
        self.F = F
        self.dims = list(dims)
        self.all_dims = list(F.dims)
        for d in self.dims:
            if d not in self.all_dims:
                raise Exception(f'dimension {d!r} not found in {tuple(self.all_dims)}')
        self.axes = [self.all_dims.index(d) for d in self.dims]
        self.values = np.asarray(F.values)
        self.shape = self.values.shape
        self.core_shape = tuple(self.shape[a] for a in self.axes)
        self.trailing = (self.axes == list(range(len(self.all_dims) - len(self.dims), len(self.all_dims))))

    def coord(self, k):
        return coord_values(self.F, self.dims[k])

    def along(self, k, v, full):
        """1-D function of core dim k -> broadcastable against the core (or full) shape."""
        n = len(self.shape) if full else len(self.dims)
        shp = [1] * n
        shp[self.axes[k] if full else k] = -1
        return np.reshape(np.asarray(v, dtype=np.float64), shp)


class _Field:
    """values + dims, the minimum core.inv_* needs from a coefficient array."""

    def __init__(self, values, dims):
        self.values = values
        self.dims = tuple(dims)


def _update(default, users, valid=None):
    """apps.__update (apps.py:2361-2375): user values override defaults unless None;
    unknown model parameters raise."""
    if valid is not None and users != default:
        for k, v in users.items():
            if k not in valid:
                raise Exception(f'mParams[\'{k}\'] is not used, valid are {valid}')
    default_cp = copy.deepcopy(default)
    for k, v in users.items():
        if v is not None:
            default_cp[k] = v
    return default_cp


def _uniform_interval(name, coord1D, value):
    if not np.isclose(np.diff(coord1D), value).all():
        raise Exception(f'coordinate {name} is non-uniform:\n{coord1D}')


def _cal_params2D(g, coords, Rearth):
    """apps.__cal_params2D (apps.py:2245-2313)."""
    dim2, dim1 = g.coord(0), g.coord(1)
    gc2, gc1 = len(dim2), len(dim1)
    del2, del1 = np.diff(dim2)[0], np.diff(dim1)[0]
    _uniform_interval(g.dims[0], dim2, del2)
    _uniform_interval(g.dims[1], dim1, del1)
    c = coords.lower()
    if c == 'lat-lon':
        del2 = np.deg2rad(del2) * Rearth
        del1 = np.deg2rad(del1) * Rearth
    elif c in ('z-lat', 'z-lon'):
        del1 = np.deg2rad(del1) * Rearth
    elif c == 'cartesian':
        pass
    else:
        raise Exception('unsupported coords for 2D case: ' + coords + ', should be [lat-lon, cartesian]')
    ratio = del1 / del2
    epsilon = np.sin(np.pi / (2.0 * gc1 + 2.0)) ** 2 + np.sin(np.pi / (2.0 * gc2 + 2.0)) ** 2
    return {
        'gc2': gc2, 'gc1': gc1, 'del2': del2, 'del1': del1, 'ratio': ratio,
        'ratioSSr': ratio ** 4.0, 'ratioSqr': ratio ** 2.0, 'ratioQtr': ratio / 4.0,
        'del1Sqr': del1 ** 2.0, 'del1Tr': del1 ** 3.0, 'del1SSr': del1 ** 4.0,
        'optArg': 2.0 / (1.0 + np.sqrt((2.0 - epsilon) * epsilon)),
        'flags': np.array([0.0, 1.0, 0.0]),
    }


def _cal_params3D(g, coords, Rearth):
    """apps.__cal_params3D (apps.py:2162-2242; the third term really uses 2*gc3+3)."""
    dim3, dim2, dim1 = g.coord(0), g.coord(1), g.coord(2)
    gc3, gc2, gc1 = len(dim3), len(dim2), len(dim1)
    del3, del2, del1 = np.diff(dim3)[0], np.diff(dim2)[0], np.diff(dim1)[0]
    _uniform_interval(g.dims[0], dim3, del3)
    _uniform_interval(g.dims[1], dim2, del2)
    _uniform_interval(g.dims[2], dim1, del1)
    c = coords.lower()
    if c == 'lat-lon':
        del2 = np.deg2rad(del2) * Rearth
        del1 = np.deg2rad(del1) * Rearth
    elif c == 'cartesian':
        pass
    else:
        raise Exception('unsupported coords for 3D case: ' + coords + ', should be in [\'lat-lon\', \'cartesian\']')
    ratio1, ratio2 = del1 / del2, del1 / del3
    epsilon = (np.sin(np.pi / (2.0 * gc1 + 2.0)) ** 2.0 + np.sin(np.pi / (2.0 * gc2 + 2.0)) ** 2.0 +
               np.sin(np.pi / (2.0 * gc3 + 3.0)) ** 2.0)
    return {
        'gc3': gc3, 'gc2': gc2, 'gc1': gc1, 'del3': del3, 'del2': del2, 'del1': del1,
        'ratio1': ratio1, 'ratio2': ratio2, 'ratio1Sqr': ratio1 ** 2.0, 'ratio2Sqr': ratio2 ** 2.0,
        'del1Sqr': del1 ** 2.0,
        'optArg': 2.0 / (1.0 + np.sqrt((2.0 - epsilon) * epsilon)),
        'flags': np.array([0.0, 1.0, 0.0]),
    }


def _mask_FS(g, iParams, icbc):
    """apps.__mask_FS (apps.py:2112-2159): maskF (undef -> -9.99e8), initS, and whether
    the reference's ``zero = maskF - maskF`` template is identically zero."""
    v = g.values
    if np.isnan(iParams['undef']):
        maskF = np.where(np.isnan(v), _undeftmp, v)
    else:
        maskF = np.where(v != iParams['undef'], v, _undeftmp)
    maskF = maskF.astype(np.result_type(v.dtype, np.float32), copy=False)
    # `zero = maskF - maskF` is identically zero unless the forcing holds an inf or an unmasked NaN
    # (then those cells turn every coefficient into NaN: a quirk the full-size path reproduces)
    zero_is_zero = bool(np.isfinite(maskF).all())
    zero = np.zeros_like(maskF) if zero_is_zero else maskF - maskF
    if icbc is None:
        return maskF, (np.zeros_like(maskF) if zero_is_zero else zero.copy()), zero, zero_is_zero
    else:
        mask = (maskF == _undeftmp)
        for k, BC in enumerate(iParams['BCs']):
            if BC != 'periodic':
                n = g.core_shape[k]
                edge = np.zeros(n, dtype=bool)
                edge[0] = edge[-1] = True       # dimV.isin([dimV[0], dimV[-1]]) for unique coordinates
                c = g.coord(k)
                edge |= np.isin(c, [c[0], c[-1]])
                mask = np.logical_or(mask, g.along(k, edge, True).astype(bool))
        initS = np.where(mask, np.asarray(icbc.values), 0).astype(zero.dtype, copy=False)
    return maskF, np.array(initS, copy=True), zero, zero_is_zero


def _coef(g, zero, shared, term):
    """A coefficient array: ``zero + term`` as the reference builds it, or -- when the
    term depends on the core dims only and ``zero`` is identically zero -- the bare
    core-shaped term, shared by every slice of the batch."""
    if shared:
        out = np.zeros(g.core_shape, dtype=np.float64) + term(False)
        return _Field(np.ascontiguousarray(out), g.dims)
    return _Field(zero + term(True), g.all_dims)


def _shift1(v):
    """``DataArray.shift({dim: 1})``: element k takes the value of k-1, NaN first."""
    out = np.empty_like(v, dtype=np.float64)
    out[0] = np.nan
    out[1:] = v[:-1]
    return out


def _coeffs_Poisson(g, coords, mParams, iParams, icbc):
    """apps.__coeffs_Poisson (apps.py:1397-1437)."""
    maskF, initS, zero, z0 = _mask_FS(g, iParams, icbc)
    c = coords.lower()
    keep = (maskF != _undeftmp)
    if c == 'lat-lon':
        lats = np.deg2rad(g.coord(0))
        cosG = np.cos(lats)
        cosH = np.cos((lats + _shift1(lats)) / 2.0)
        A = _coef(g, zero, z0, lambda full: g.along(0, cosH, full))
        B = None
        C = _coef(g, zero, z0, lambda full: g.along(0, 1.0 / cosG, full))
        Fm = np.where(keep, maskF * g.along(0, cosG, True), _undeftmp)
    elif c == 'z-lat':
        cosG = np.cos(np.deg2rad(g.coord(1)))
        A = _coef(g, zero, z0, lambda full: 1.0)
        B = None
        C = _coef(g, zero, z0, lambda full: 1.0)
        Fm = np.where(keep, maskF * g.along(1, cosG, True), _undeftmp)
    elif c in ('z-lon', 'cartesian'):
        A = _coef(g, zero, z0, lambda full: 1.0)
        B = None
        C = _coef(g, zero, z0, lambda full: 1.0)
        Fm = np.where(keep, maskF, _undeftmp)
    else:
        raise Exception('unsupported coords ' + coords + ', should be in [lat-lon, z-lat, z-lon, cartesian]')
    if B is None:
        B = _Field(zero, g.all_dims) if not z0 else _Field(np.zeros(g.core_shape), g.dims)
    return maskF, Fm, initS, (A, B, C)


def _poisson_device_front(F, dims, coords, icbc, mParams, iParams):
    """invert_Poisson through the device-side front end, or None when the reference-shaped host
    path has to be taken (icbc given, core dims not trailing, non-float64 input, lexicographic
    ordering / colour engine requested, odd nx with periodic-x, non-finite forcing values ...).
    Follows apps.__template (apps.py:1324-1394) step for step; the arrays it would build are
    described to the library by three vectors along the first core dim instead."""
    if icbc is not None or len(dims) != 2:
        return None
    ip = _update(default_iParams, iParams)
    if ip.get('ordering', 'colour') not in ('colour', 'color', 'redblack', 'red-black') or \
            ip.get('engine', 'auto') == 'colour' or core.solvers is not _device_solvers:
        return None
    mp = _update(default_mParams, mParams, ['g', 'Omega', 'Rearth'])
    g = _Grid(F, dims)
    if not g.trailing or g.values.dtype not in (np.float64, np.float32):     # (float32: xinv_opts.io_f32)
        return None
    c = coords.lower()
    ny = g.core_shape[0]
    if c == 'lat-lon':
        lats = np.deg2rad(g.coord(0))
        cosG = np.cos(lats)
        A_rows, C_rows, scale = np.cos((lats + _shift1(lats)) / 2.0), 1.0 / cosG, cosG
    elif c == 'z-lat':
        return None                              # the forcing scale runs along the second core dim
    elif c in ('z-lon', 'cartesian'):
        A_rows, C_rows, scale = np.ones(ny), np.ones(ny), None
    else:
        return None                              # let the host path raise the reference's exception
    ps = _cal_params2D(g, coords, mp['Rearth'])
    ip = _update(ps, ip)
    if ip['debug']:
        _print_params(ip)
    try:
        S, flags, stats = _device_solvers.solve_standard_2D_rows(
            g.values, A_rows, C_rows, scale, ip['undef'], ip['undef'], ip['BCs'][0], ip['BCs'][1], ip['del1Sqr'],
            ip['ratioQtr'], ip['ratioSqr'], ip['optArg'], _undeftmp, ip['flags'], ip['mxLoop'], ip['tolerance'],
            ctx=ip.get('ctx'), devices=ip.get('devices'), accel=ip.get('accel'))
    except XinvError as e:
        if e.code == E_UNSUPPORTED:              # not a problem for the fused engine
            return None
        raise
    _, noncore, _ = core._layout(F, dims)
    core._report(ip, core._slice_labels(F, noncore), flags)
    _report_flags(iParams, flags, stats)
    return wrap_like(F, S, name='inverted')


def _user_field(g, zero, X, name):
    """``zero + X`` for a user-supplied coefficient (scalar, core-shaped or full)."""
    if hasattr(X, 'dims'):
        xd = list(X.dims)
        xv = np.asarray(X.values, dtype=np.float64)
        if xd == g.dims:
            return _Field(np.ascontiguousarray(xv), g.dims)
        if xd == g.all_dims:
            return _Field(zero + xv, g.all_dims)
        if len(xd) == 1 and xd[0] in g.dims:    # 1-D profile along one core dim
            k = g.dims.index(xd[0])
            return _Field(np.ascontiguousarray(np.zeros(g.core_shape) + g.along(k, xv, False)), g.dims)
        raise Exception(f"mParams['{name}'] has dims {tuple(xd)}; expected {tuple(g.dims)} or {tuple(g.all_dims)}")
    return _Field(np.zeros(g.core_shape) + float(X), g.dims)


def _coeffs_Eliassen(g, coords, mParams, iParams, icbc):
    """apps.__coeffs_Eliassen (apps.py:1582-1606)."""
    maskF, initS, zero, z0 = _mask_FS(g, iParams, icbc)
    if coords.lower() not in ('z-lat', 'cartesian'):
        raise Exception('unsupported coords ' + coords + ', should be in [z-lat, cartesian]')
    A = _user_field(g, zero, mParams['A'], 'A')
    B = _user_field(g, zero, mParams['B'], 'B')
    C = _user_field(g, zero, mParams['C'], 'C')
    Fm = np.where(maskF != _undeftmp, maskF, _undeftmp)
    return maskF, Fm, initS, (A, B, C)


def _eliassen_device_front(F, dims, coords, icbc, mParams, iParams):
    """invert_Eliassen through the dense device front end (xinv_std2d_front), or None when the reference-shaped host
    path has to be taken (icbc given, core dims not trailing, lexicographic ordering, non-float64 input, a coefficient
    whose dims are neither the core dims, all dims, one core dim nor a scalar).  Follows apps.__template
    (apps.py:1324-1394) with apps.__coeffs_Eliassen (:1582-1606): the coefficient fields are handed over as the user
    holds them (core-shaped ones shared by the whole batch), masks / zero guess / de-masking happen on the device."""
    if icbc is not None or len(dims) != 2:
        return None
    ip = _update(default_iParams, iParams)
    if ip.get('ordering', 'colour') not in ('colour', 'color', 'redblack', 'red-black') or core.solvers is not _device_solvers:
        return None
    mp = _update(default_mParams, mParams, ['A', 'B', 'C', 'g', 'Omega', 'Rearth'])
    g = _Grid(F, dims)
    if not g.trailing or g.values.dtype != np.float64 or coords.lower() not in ('z-lat', 'cartesian'):
        return None

    def field(X):                                # _user_field without the full-size zero template
        if hasattr(X, 'dims'):
            xd, xv = list(X.dims), np.ascontiguousarray(np.asarray(X.values, dtype=np.float64))
            if xd == g.dims or (xd == g.all_dims and xd != g.dims):
                return xv
            if len(xd) == 1 and xd[0] in g.dims:
                return np.ascontiguousarray(np.zeros(g.core_shape) + g.along(g.dims.index(xd[0]), xv, False))
            return None
        return np.zeros(g.core_shape) + float(X)

    fields = [field(mp[k]) for k in ('A', 'B', 'C')]
    if any(f is None for f in fields):
        return None
    ps = _cal_params2D(g, coords, mp['Rearth'])
    ip = _update(ps, ip)
    if ip['debug']:
        _print_params(ip)
    try:
        S, flags, stats = _device_solvers.solve_standard_2D_front(
            g.values, fields[0], fields[1], fields[2], ip['undef'], ip['undef'], ip['BCs'][0], ip['BCs'][1], ip['del1Sqr'],
            ip['ratioQtr'], ip['ratioSqr'], ip['optArg'], _undeftmp, ip['flags'], ip['mxLoop'], ip['tolerance'],
            engine=ip.get('engine', 'auto'), ctx=ip.get('ctx'), devices=ip.get('devices'), accel=ip.get('accel'))
    except XinvError as e:
        if e.code == E_UNSUPPORTED:              # e.g. a non-finite unmasked forcing value
            return None
        raise
    _, noncore, _ = core._layout(F, dims)
    core._report(ip, core._slice_labels(F, noncore), flags)
    _report_flags(iParams, flags, stats)
    return wrap_like(F, S, name='inverted')


def _rows_GillMatsuno(g, coords, mParams):
    """The coefficients of apps.__coeffs_GillMatsuno (apps.py:1609-1657) as functions of the first core
    dim: A, C, D, E (vectors) and F (a constant); the forcing is used as it is (g_mode 0)."""
    Phi, epsilon = mParams['Phi'], mParams['epsilon']
    f0, beta = mParams['f0'], mParams['beta']
    Omega, Rearth = mParams['Omega'], mParams['Rearth']
    ydef = np.asarray(g.coord(0), dtype=np.float64)
    c = coords.lower()
    if c == 'lat-lon':
        lats = np.deg2rad(ydef)
        cosL = np.cos(lats)
        f = 2.0 * Omega * np.sin(lats)
        c1 = epsilon / (epsilon ** 2. + f ** 2.)
        c2 = f / (epsilon ** 2. + f ** 2.)
        deg2m = Rearth / 180. * np.pi
        dc1 = np.gradient(c1, ydef, edge_order=1)       # c1.differentiate(lat), per degree
        dc2 = np.gradient(c2, ydef, edge_order=1)
        tA = c1 * Phi
        tC = c1 * Phi / cosL ** 2.
        tD = Phi * (dc1 / deg2m + c1 * np.tan(lats) / Rearth)
        tE = - Phi * dc2 / deg2m / cosL
    elif c == 'cartesian':
        f = f0 + beta * ydef
        c1 = epsilon / (epsilon ** 2. + f ** 2.)
        c2 = f / (epsilon ** 2. + f ** 2.)
        tA = c1 * Phi
        tC = c1 * Phi
        tD = Phi * np.gradient(c1, ydef, edge_order=1)
        tE = - Phi * np.gradient(c2, ydef, edge_order=1)
    else:
        raise Exception('unsupported coords ' + coords + ', should be in [lat-lon, cartesian]')
    return dict(A=tA, C=tC, D=tD, E=tE, F=-epsilon, g_mode=0, g_p1=1.0, g_p2=1.0)


def _coeffs_GillMatsuno(g, coords, mParams, iParams, icbc):
    """apps.__coeffs_GillMatsuno (apps.py:1609-1657)."""
    r = _rows_GillMatsuno(g, coords, mParams)
    maskF, initS, zero, z0 = _mask_FS(g, iParams, icbc)
    A = _coef(g, zero, z0, lambda full: g.along(0, r['A'], full))
    B = _Field(zero, g.all_dims) if not z0 else _Field(np.zeros(g.core_shape), g.dims)
    C = _coef(g, zero, z0, lambda full: g.along(0, r['C'], full))
    D = _coef(g, zero, z0, lambda full: g.along(0, r['D'], full))
    E = _coef(g, zero, z0, lambda full: g.along(0, r['E'], full))
    F = _coef(g, zero, z0, lambda full: r['F'])
    G = np.where(maskF != _undeftmp, maskF, _undeftmp)
    return maskF, G, initS, (A, B, C, D, E, F)


def _rows_Stommel(g, coords, mParams):
    """apps.__coeffs_Stommel (apps.py:1712-1748) along the first core dim: A, C, E (B = D = F = 0) and the
    forcing transform G = -curl / D / rho0 (g_mode 1)."""
    beta, R, depth, rho0 = mParams['beta'], mParams['R'], mParams['D'], mParams['rho0']
    Rearth, Omega = mParams['Rearth'], mParams['Omega']
    c = coords.lower()
    if c == 'lat-lon':
        cosL = np.cos(np.deg2rad(g.coord(0)))
        tA, tC, tE = - R / depth, - R / depth / cosL ** 2., - 2. * Omega / Rearth
    elif c == 'cartesian':
        tA, tC, tE = - R / depth, - R / depth, - beta
    else:
        raise Exception('unsupported coords ' + coords + ', should be in [lat-lon, z-lat, z-lon, cartesian]')
    return dict(A=tA, C=tC, D=0.0, E=tE, F=0.0, g_mode=1, g_p1=depth, g_p2=rho0)


def _coeffs_Stommel(g, coords, mParams, iParams, icbc):
    """apps.__coeffs_Stommel (apps.py:1712-1748)."""
    r = _rows_Stommel(g, coords, mParams)
    depth, rho0 = mParams['D'], mParams['rho0']
    maskF, initS, zero, z0 = _mask_FS(g, iParams, icbc)
    vec = lambda v: (lambda full: g.along(0, v, full)) if np.ndim(v) else (lambda full: v)
    A = _coef(g, zero, z0, vec(r['A']))
    C = _coef(g, zero, z0, vec(r['C']))
    E = _coef(g, zero, z0, vec(r['E']))
    zf = _Field(zero, g.all_dims) if not z0 else _Field(np.zeros(g.core_shape), g.dims)
    B, D, F = zf, zf, zf
    G = np.where(maskF != _undeftmp, -maskF / depth / rho0, _undeftmp)
    return maskF, G, initS, (A, B, C, D, E, F)


def _general_device_front(rows_func, valid, F, dims, coords, icbc, mParams, iParams):
    """invert_GillMatsuno / invert_Stommel through the device-side front end (``xinv_gen2d_rows``), or
    None when the reference-shaped host path has to be taken (see _poisson_device_front)."""
    if icbc is not None or len(dims) != 2:
        return None
    ip = _update(default_iParams, iParams)
    if ip.get('ordering', 'colour') not in ('colour', 'color', 'redblack', 'red-black') or \
            ip.get('engine', 'auto') == 'colour' or core.solvers is not _device_solvers:
        return None
    mp = _update(default_mParams, mParams, valid)
    g = _Grid(F, dims)
    if not g.trailing or g.values.dtype not in (np.float64, np.float32) or coords.lower() not in ('lat-lon', 'cartesian'):
        return None
    r = rows_func(g, coords, mp)
    ny = g.core_shape[0]
    rows = np.stack([np.broadcast_to(np.asarray(r[k], dtype=np.float64), (ny,)) for k in 'ACDEF'])
    ps = _cal_params2D(g, coords, mp['Rearth'])
    ip = _update(ps, ip)
    if ip['debug']:
        _print_params(ip)
    try:
        S, flags, stats = _device_solvers.solve_general_2D_rows(
            g.values, rows, r['g_mode'], r['g_p1'], r['g_p2'], ip['undef'], ip['undef'], ip['BCs'][0], ip['BCs'][1],
            ip['del1'], ip['del1Sqr'], ip['ratio'], ip['ratioQtr'], ip['ratioSqr'], ip['optArg'], _undeftmp,
            ip['flags'], ip['mxLoop'], ip['tolerance'], ctx=ip.get('ctx'), devices=ip.get('devices'), accel=ip.get('accel'))
    except XinvError as e:
        if e.code == E_UNSUPPORTED:              # not a problem for the fused engine
            return None
        raise
    _, noncore, _ = core._layout(F, dims)
    core._report(ip, core._slice_labels(F, noncore), flags)
    _report_flags(iParams, flags, stats)
    return wrap_like(F, S, name='inverted')


def _omega_device_front(F, dims, coords, icbc, mParams, iParams):
    """invert_omega through the device-side front end (``xinv_std3d_rows``), or None when the reference-shaped
    host path has to be taken (see _poisson_device_front).  A, the factor of B, the divisor of C and the forcing
    scale are vectors along the second core dim; N2 (scalar, profile, volume or full array) is handed over as it
    is and read through strides."""
    if icbc is not None or len(dims) != 3:
        return None
    ip = _update(default_iParams, iParams)
    if ip.get('ordering', 'colour') not in ('colour', 'color', 'redblack', 'red-black') or \
            ip.get('engine', 'auto') == 'colour' or core.solvers is not _device_solvers:
        return None
    mp = _update(default_mParams, mParams, ['f0', 'beta', 'N2', 'g', 'Omega', 'Rearth'])
    g = _Grid(F, dims)
    if not g.trailing or g.values.dtype not in (np.float64, np.float32):
        return None
    c = coords.lower()
    nz, ny, nx = g.core_shape
    ydef = np.asarray(g.coord(1), dtype=np.float64)
    if c == 'lat-lon':
        lats = np.deg2rad(ydef)
        cosG = np.cos(lats)
        f = 2. * mp['Omega'] * np.sin(lats)
        rows = np.stack([f ** 2 * cosG, np.cos((lats + _shift1(lats)) / 2.), cosG, cosG])
    elif c == 'cartesian':
        f = mp['f0'] + mp['beta'] * ydef
        rows = np.stack([f ** 2., np.ones(ny), np.ones(ny), np.ones(ny)])
    else:
        return None                              # let the host path raise the reference's exception
    N2 = mp['N2']
    if hasattr(N2, 'dims'):
        nd = list(N2.dims)
        nv = np.ascontiguousarray(np.asarray(N2.values, dtype=np.float64))
        core_strides = {g.dims[0]: ny * nx, g.dims[1]: nx, g.dims[2]: 1}
        if nd == g.all_dims and len(g.all_dims) > 3:
            strides = (nz * ny * nx, ny * nx, nx, 1)
        elif nd == g.dims:
            strides = (0, ny * nx, nx, 1)
        elif len(nd) == 1 and nd[0] in g.dims:
            k = g.dims.index(nd[0])
            strides = tuple([0] + [1 if m == k else 0 for m in range(3)])
        else:
            return None
        del core_strides
    else:
        nv, strides = np.array([float(N2)]), (0, 0, 0, 0)
    ps = _cal_params3D(g, coords, mp['Rearth'])
    ip = _update(ps, ip)
    if ip['debug']:
        _print_params(ip)
    try:
        S, flags, stats = _device_solvers.solve_standard_3D_rows(
            g.values, rows, nv, strides, ip['undef'], ip['undef'], ip['BCs'][0], ip['BCs'][1], ip['BCs'][2], ip['del1Sqr'],
            ip['ratio2Sqr'], ip['ratio1Sqr'], ip['optArg'], _undeftmp, ip['flags'], ip['mxLoop'], ip['tolerance'],
            ctx=ip.get('ctx'), devices=ip.get('devices'), accel=ip.get('accel'))
    except XinvError as e:
        if e.code == E_UNSUPPORTED:              # not a problem for the fused engine
            return None
        raise
    _, noncore, _ = core._layout(F, dims)
    core._report(ip, core._slice_labels(F, noncore), flags)
    _report_flags(iParams, flags, stats)
    return wrap_like(F, S, name='inverted')


def _coeffs_omega(g, coords, mParams, iParams, icbc):
    """apps.__coeffs_omega (apps.py:2016-2052).  ``N2`` may be a scalar, a 1-D
    profile along a core dim, a core-shaped or a full-shaped array."""
    f0, beta, N2, Omega = mParams['f0'], mParams['beta'], mParams['N2'], mParams['Omega']
    maskF, initS, zero, z0 = _mask_FS(g, iParams, icbc)
    ydef = np.asarray(g.coord(1), dtype=np.float64)

    n2_full = hasattr(N2, 'dims') and list(N2.dims) == g.all_dims and list(N2.dims) != g.dims
    shared = z0 and not n2_full

    def n2(full):
        if not hasattr(N2, 'dims'):
            return float(N2)
        nd = list(N2.dims)
        nv = np.asarray(N2.values, dtype=np.float64)
        if nd == g.all_dims and full:
            return nv
        if nd == g.dims:
            if not full:
                return nv
            shp = [1] * len(g.shape)
            for k, a in enumerate(g.axes):
                shp[a] = g.core_shape[k]
            return nv.reshape(shp) if g.trailing else np.moveaxis(
                nv.reshape([1] * (len(g.shape) - len(g.dims)) + list(g.core_shape)),
                range(len(g.shape) - len(g.dims), len(g.shape)), g.axes)
        if len(nd) == 1 and nd[0] in g.dims:
            return g.along(g.dims.index(nd[0]), nv, full)
        raise Exception(f"mParams['N2'] has dims {tuple(nd)}; expected a subset of {tuple(g.all_dims)}")

    c = coords.lower()
    if c == 'lat-lon':
        lats = np.deg2rad(ydef)
        cosH = np.cos((lats + _shift1(lats)) / 2.)
        cosG = np.cos(lats)
        f = 2. * Omega * np.sin(lats)
        A = _coef(g, zero, shared, lambda full: g.along(1, f ** 2 * cosG, full))
        B = _coef(g, zero, shared, lambda full: n2(full) * g.along(1, cosH, full))
        C = _coef(g, zero, shared, lambda full: n2(full) / g.along(1, cosG, full))
        Fm = np.where(maskF != _undeftmp, maskF * g.along(1, cosG, True), _undeftmp)
    elif c == 'cartesian':
        f = f0 + beta * ydef
        A = _coef(g, zero, shared, lambda full: g.along(1, f ** 2., full))
        B = _coef(g, zero, shared, lambda full: n2(full) + 0.0)
        C = _coef(g, zero, shared, lambda full: n2(full) + 0.0)
        Fm = np.where(maskF != _undeftmp, maskF, _undeftmp)
    else:
        raise Exception('unsupported coords ' + coords + ', should be in [lat-lon, cartesian]')
    return maskF, Fm, initS, (A, B, C)


def _report_flags(user_iParams, flags, stats=None):
    """Per-slice flags [batch, 3] (and the library's statistics of the call) go back to the caller's own
    iParams dict (an addition to the reference's interface, the same on every path) -- never into the
    module-level defaults."""
    if isinstance(user_iParams, dict) and user_iParams is not default_iParams:
        user_iParams['flags_all'] = flags
        if stats is not None:
            user_iParams['stats'] = stats


def _print_params(iParams):
    for k in sorted(iParams):
        if k not in ('flags_all',):
            print(f'{k:10s}: {iParams[k]}')


def _template(coef_func, inv_func, dimLen, F, dims, coords='lat-lon', icbc=None,
              validParams=[], mParams=default_mParams, iParams=default_iParams):
    """apps.__template (apps.py:1324-1394)."""
    if len(dims) != dimLen:
        raise Exception('{0:2d} dimensional forcing are needed'.format(dimLen))

    user_iParams = iParams
    iParams = _update(default_iParams, iParams)
    mParams = _update(default_mParams, mParams, validParams)
    g = _Grid(F, dims)

    ######  1. calculating the coefficients  ######
    maskF, forcing, initS, coeffs = coef_func(g, coords, mParams, iParams, icbc)

    ######  2. calculating the parameters  ######
    if dimLen == 2:
        ps = _cal_params2D(g, coords, mParams['Rearth'])
    elif dimLen == 3:
        ps = _cal_params3D(g, coords, mParams['Rearth'])
    else:
        raise Exception('dimension length should be one of [2, 3]')
    iParams = _update(ps, iParams)
    if iParams['debug']:
        _print_params(iParams)

    ######  3. inverting the solution  ######
    S = wrap_like(F, np.ascontiguousarray(initS, dtype=np.float64))
    Ff = wrap_like(F, forcing)
    inv_func(*coeffs, Ff, S, dims, iParams)
    if 'flags_all' in iParams:
        _report_flags(user_iParams, iParams['flags_all'], iParams.get('stats'))

    ######  4. properly de-masking  ######
    out = np.asarray(S.values)                   # our own array (built from initS), free to modify
    if icbc is None:
        out[maskF == _undeftmp] = iParams['undef']
    if np.asarray(F.values).dtype == np.float32:
        out = out.astype(np.float32)
    return wrap_like(F, out, name='inverted')
