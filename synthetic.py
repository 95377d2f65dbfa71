"""Seeded synthetic problems of the BASELINE.json configs (SURVEY.md 8d recipes), ndarray level.

Shared by bench.py, scripts/ and tests/cases.py (which re-exports them): the coefficient recipes
restate the reference's builders in numpy -- apps.py:1401-1409 (lat-lon Poisson), :2025-2036
(omega), :1638-1651 (Gill-Matsuno on a beta plane), :2245-2313 / :2162-2242 (grid parameters).
"""
import numpy as np

UNDEF = -9.99e8
REARTH = 6371200.0
OMEGA = 7.292e-5


def params2d(ny, nx, del2, del1):
    """numpy restatement of apps.__cal_params2D (apps.py:2282-2291)."""
    ratio = del1 / del2
    eps = np.sin(np.pi / (2.0 * nx + 2.0)) ** 2 + np.sin(np.pi / (2.0 * ny + 2.0)) ** 2
    return dict(gc2=ny, gc1=nx, del2=del2, del1=del1, ratio=ratio, ratioSqr=ratio ** 2.0,
                ratioQtr=ratio / 4.0, del1Sqr=del1 ** 2.0,
                optArg=2.0 / (1.0 + np.sqrt((2.0 - eps) * eps)))


def params3d(nz, ny, nx, del3, del2, del1):
    """numpy restatement of apps.__cal_params3D (apps.py:2162-2242; the third term really uses 2*gc3+3)."""
    eps = (np.sin(np.pi / (2.0 * nx + 2.0)) ** 2.0 + np.sin(np.pi / (2.0 * ny + 2.0)) ** 2.0 +
           np.sin(np.pi / (2.0 * nz + 3.0)) ** 2.0)
    return dict(gc3=nz, gc2=ny, gc1=nx, del3=del3, del2=del2, del1=del1, del1Sqr=del1 ** 2.0,
                ratio1Sqr=(del1 / del2) ** 2.0, ratio2Sqr=(del1 / del3) ** 2.0,
                optArg=2.0 / (1.0 + np.sqrt((2.0 - eps) * eps)))


def latlon_grid(ny, nx):
    dlat, dlon = 180.0 / ny, 360.0 / nx
    return -90.0 + dlat / 2 + dlat * np.arange(ny), dlon * np.arange(nx)


def poisson_latlon(ny, nx, land=True, noise=1e-6, seed=0, batch=None, phase=0.0):
    """lat-lon Poisson problem: A=cosH, C=1/cosG, F=zeta*cosG with an optional
    land mask (apps.py:1401-1409); zeta is the SURVEY.md 8d formula."""
    dlat, dlon = 180.0 / ny, 360.0 / nx
    lat, lon = latlon_grid(ny, nx)
    lats = np.deg2rad(lat)
    cosG = np.cos(lats)
    latm = np.empty(ny)
    latm[0] = np.nan
    latm[1:] = lats[:-1]
    cosH = np.cos((lats + latm) / 2.0)
    lam = np.deg2rad(lon)[None, :]
    phi = lats[:, None]
    rng = np.random.default_rng(seed)
    shape = (ny, nx) if batch is None else (batch, ny, nx)
    ph = phase if batch is None else (2 * np.pi * np.arange(batch) / batch)[:, None, None]
    zeta = 1e-5 * np.sin(3 * lam + ph) * np.cos(phi) ** 2 * np.sin(2 * phi)
    zeta = np.broadcast_to(zeta, shape) + noise * rng.standard_normal(shape)
    zeta = np.ascontiguousarray(zeta)
    F = zeta * cosG[:, None]
    if land:
        mask = np.sin(5 * lam) * np.cos(3 * phi) > 0.6
        F[..., mask] = UNDEF
    A = np.ascontiguousarray(np.broadcast_to(cosH[:, None], (ny, nx)))
    C = np.ascontiguousarray(np.broadcast_to(1.0 / cosG[:, None], (ny, nx)))
    p = params2d(ny, nx, np.deg2rad(dlat) * REARTH, np.deg2rad(dlon) * REARTH)
    return dict(A=A, C=C, F=F, p=p, S0=np.zeros(shape))


def poisson_latlon_user(ny, nx, land=True, noise=1e-6, seed=0, phase=0.0):
    """The same problem as poisson_latlon the way a user of invert_Poisson holds it: the raw
    vorticity (NaN on land) and the lat / lon coordinates."""
    lat, lon = latlon_grid(ny, nx)
    lam, phi = np.deg2rad(lon)[None, :], np.deg2rad(lat)[:, None]
    rng = np.random.default_rng(seed)
    zeta = 1e-5 * np.sin(3 * lam + phase) * np.cos(phi) ** 2 * np.sin(2 * phi) + noise * rng.standard_normal((ny, nx))
    if land:
        zeta[np.sin(5 * lam) * np.cos(3 * phi) > 0.6] = np.nan
    return zeta, lat, lon


def omega_latlon(nz, ny, nx, seed=1, n2="3d", dlev=-2500.0, lat0=None, dlat=None, dlon=None):
    """invert_omega on a lat-lon grid (apps.py:2025-2036): A = f^2 cosG, B = N2 cosH, C = N2 / cosG,
    F = forcing * cosG.  n2: '3d' (a different value in every cell, BASELINE configs[2]: "variable
    coeffs") or '1d' (a profile along the levels, as in the reference's notebook 11)."""
    rng = np.random.default_rng(seed)
    if lat0 is None:
        lat, lon = latlon_grid(ny, nx)
        dlat, dlon = 180.0 / ny, 360.0 / nx
    else:
        lat, lon = lat0 + dlat * np.arange(ny), 140.0 + dlon * np.arange(nx)
    lats = np.deg2rad(lat)
    cosG = np.cos(lats)
    latm = np.empty(ny)
    latm[0] = np.nan
    latm[1:] = lats[:-1]
    cosH = np.cos((lats + latm) / 2.0)
    f = 2.0 * OMEGA * np.sin(lats)
    shape = (nz, ny, nx)
    if n2 == "3d":
        N2 = 1e-6 * (1.0 + 0.5 * rng.random(shape))
    else:
        N2 = np.broadcast_to((1e-6 * (1.0 + 0.5 * rng.random(nz)))[:, None, None], shape)
    A = np.ascontiguousarray(np.broadcast_to((f ** 2 * cosG)[None, :, None], shape))
    B = np.ascontiguousarray(N2 * cosH[None, :, None])
    C = np.ascontiguousarray(N2 / cosG[None, :, None])
    F = 1e-17 * rng.standard_normal(shape) * cosG[None, :, None]
    p = params3d(nz, ny, nx, dlev, np.deg2rad(dlat) * REARTH, np.deg2rad(dlon) * REARTH)
    return dict(A=A, B=B, C=C, F=F, S0=np.zeros(shape), p=p)


def gill_matsuno_beta(ny, nx):
    """invert_GillMatsuno on a cartesian beta plane (apps.py:1638-1651): f = beta y,
    c1 = eps / (eps^2 + f^2), c2 = f / (eps^2 + f^2), A = C = c1 Phi, D = Phi dc1/dy, E = -Phi dc2/dy,
    F = -eps, G = Q."""
    beta, eps, Phi = 2e-11, 1e-5, 5000.0
    y, x = np.linspace(-5e6, 5e6, ny), np.linspace(0, 4e7, nx, endpoint=False)
    f = beta * y
    c1, c2 = eps / (eps ** 2 + f ** 2), f / (eps ** 2 + f ** 2)
    row = lambda v: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64).reshape(-1, 1) if np.ndim(v) else v,
                                                         (ny, nx)) + 0.0)
    yy, xx = np.meshgrid(y, x, indexing="ij")
    G = 0.05 * np.exp(-((yy / 1e6) ** 2 + ((xx - 2e7) / 2e6) ** 2))
    c = dict(A=row(c1 * Phi), B=None, C=row(c1 * Phi), D=row(Phi * np.gradient(c1, y, edge_order=1)),
             E=row(-Phi * np.gradient(c2, y, edge_order=1)), F=row(-eps), G=G, S0=np.zeros((ny, nx)))
    c["p"] = params2d(ny, nx, y[1] - y[0], x[1] - x[0])
    return c
